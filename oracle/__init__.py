"""CPU oracle for iago_b200 — TEST INFRASTRUCTURE ONLY.

Nothing under this directory is product code.  Only `tests/`, `__graft_entry__.smoke()` and
`bench.py`'s `cpu_baseline` / `--impl reference` legs may import it; the product package
`iago_b200` never does (tests/test_no_oracle_in_product.py enforces that).
"""
