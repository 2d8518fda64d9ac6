#!/usr/bin/env python
"""Write profiles/ncu_traffic.json — the measured per-launch figures bench.py quotes — from the committed ncu captures:

    python profiles/ncu_traffic.py gpurun_out/r02b_rollout_philox.ncu-rep gpurun_out/r02b_rollout_forced.ncu-rep

rollout_pair_kernel: dram__bytes_read/write per launch (the bench line's roofline.traffic);
rollout_pair_kernel_forced: warp instructions the ALU pipe executed in the rules-only launch (x 32 lanes / plies of the launch = the
ALU lane-ops per ply the kernel really issues for move generation + flips: roofline_movegen.frac_as_implemented)."""
import csv, datetime, json, os, subprocess, sys

HERE = os.path.dirname(os.path.abspath(__file__))


def row(path, want):
    out = subprocess.run(["ncu", "-i", path, "--page", "raw", "--csv"], capture_output=True, text=True).stdout
    rows = list(csv.reader(out.splitlines()))
    hdr, units = rows[0], rows[1]
    for r in rows[2:]:
        if want in r[hdr.index("Kernel Name")]:
            def get(m):
                v, u = float(r[hdr.index(m)].replace(",", "")), units[hdr.index(m)]
                return v * {"Kbyte": 1e3, "Mbyte": 1e6, "Gbyte": 1e9}.get(u, 1.0)
            return get
    raise SystemExit(f"no kernel matching {want} in {path}")


def main():
    philox, forced = sys.argv[1], sys.argv[2]
    today = datetime.date.today().isoformat()
    cmd = "python bench.py --steps 3 --warmup 3 --no-cpu --sections rollout"
    g = row(philox, "rollout_pair_kernel")
    rd, wr = int(g("dram__bytes_read.sum")), int(g("dram__bytes_write.sum"))
    out = {"rollout_pair_kernel": {
        "dram_bytes_per_launch": rd + wr, "dram_bytes_read": rd, "dram_bytes_write": wr, "games_per_launch": 65536,
        "alu_pipe_warp_inst_per_launch": int(g("sm__inst_executed_pipe_alu.sum")), "warp_inst_per_launch": int(g("sm__inst_executed.sum")),
        "alu_pipe_pct": g("sm__inst_executed_pipe_alu.avg.pct_of_peak_sustained_active"),
        "issue_pct": g("smsp__issue_active.avg.pct_of_peak_sustained_active"), "kernel_us": g("gpu__time_duration.sum"),
        "source": f"{os.path.basename(philox)} (ncu --set full --clock-control none, profiles/run_ncu_r02b.sh)", "captured": today, "command": cmd}}
    f = row(forced, "rollout_pair_kernel")
    out["rollout_pair_kernel_forced"] = {
        "games_per_launch": 65536, "alu_pipe_warp_inst_per_launch": int(f("sm__inst_executed_pipe_alu.sum")),
        "fma_pipe_warp_inst_per_launch": int(f("sm__inst_executed_pipe_fma.sum")), "warp_inst_per_launch": int(f("sm__inst_executed.sum")),
        "alu_pipe_pct": f("sm__inst_executed_pipe_alu.avg.pct_of_peak_sustained_active"),
        "issue_pct": f("smsp__issue_active.avg.pct_of_peak_sustained_active"), "kernel_us": f("gpu__time_duration.sum"),
        "source": f"{os.path.basename(forced)} (the same command, FORCED instantiation)", "captured": today, "command": cmd}
    json.dump(out, open(os.path.join(HERE, "ncu_traffic.json"), "w"), indent=1)
    print(json.dumps(out, indent=1))


if __name__ == "__main__":
    main()
