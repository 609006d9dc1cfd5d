"""TEST INFRASTRUCTURE ONLY -- CPU restatement of the reference's solve path.

Nothing under `neptune_mip_b200/` imports this package.  Allowed importers: `tests/`,
`__graft_entry__.smoke()` (as the checker) and `bench.py`'s `cpu_baseline` / `--impl reference`
legs.  See DESIGN.md section "Oracle".

Parity pinning: `oracle.model`, `oracle.checkers`, `oracle.efttc` are checked (tests/test_oracle_*.py)
against (i) the reference's own golden outputs (`output-mip.json`, the six Alibaba
`output_*_case0.json`), committed as fixtures under `tests/golden/` by `oracle/make_golden.py`,
and (ii) when /root/reference is present, against the unmodified reference code imported
through `oracle.refshim` (HiGHS stands in for the un-vendored OR-Tools/SCIP wheel,
`requirements.txt:8`).
"""
