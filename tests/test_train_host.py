"""Host-side logic of the training loop (no GPU): the pool / snapshot / early-stop rule of src/train_rl.py:67-81."""
import numpy as np


def reference_rule(rates, models=1):
    """Literal restatement of src/train_rl.py:71-81 for a sequence of win rates."""
    cnt, snaps = 0, []
    for i, rate in enumerate(rates):
        if rate > 0.5:
            cnt += 1
        if cnt > 4 * np.sqrt(models) and rate > 0.6:
            snaps.append((i, models))
            models += 1
            cnt = 0
        if rate < 0.2:
            return snaps, i
        if models > 20:
            return snaps, i
    return snaps, None


def test_pool_schedule_matches_reference_rule():
    from iago_b200.train_rl import PoolSchedule
    rng = np.random.default_rng(0)
    for trial in range(50):
        rates = np.clip(rng.normal(0.6, 0.15, size=400), 0, 1)
        if trial % 5 == 0:
            rates[rng.integers(50, 400)] = 0.1
        want_snaps, want_stop = reference_rule(rates)
        s = PoolSchedule(1)
        snaps, stopped = [], None
        for i, r in enumerate(rates):
            if not s.running():
                stopped = i - 1
                break
            snap, stop = s.step(float(r))
            if snap is not None:
                snaps.append((i, snap))
            if stop:
                stopped = i
                break
        assert snaps == want_snaps
        assert stopped == want_stop
