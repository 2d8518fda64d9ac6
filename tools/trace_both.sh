IAGO_NVCC_EXTRA=-DIAGO_TRUNK_TRACE python -m iago_b200.build --force > /dev/null 2>&1
python tools/trace_trunk.py > gpurun_out/r02_trunk_trace_v3_pair.log 2>&1
IAGO_TRUNK_CG1=1 python tools/trace_trunk.py > gpurun_out/r02_trunk_trace_v3_cg1.log 2>&1
