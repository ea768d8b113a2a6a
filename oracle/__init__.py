"""Oracle = TEST INFRASTRUCTURE ONLY.

CPU restatement (`restatement.py`) of the reference's prompted 3D shifted-window
attention block plus a loader for the live reference (`ref_loader.py`, only usable
where /root/reference is mounted).  Nothing under the product package may import
this directory: only `tests/`, `__graft_entry__.smoke()` and the `cpu_baseline` /
`--impl reference` legs of `bench.py` do, and there only as the checker.

Parity pinning: the reference ships no tests or golden vectors, so
the restatement is pinned against outputs of the reference itself, executed in the
build container by `oracle/gen_golden.py` and committed under `tests/golden/`.
"""
