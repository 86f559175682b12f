"""CPU restatement of the Show-o hot path (test infrastructure only).

`showo_oracle` / `magvit_oracle` restate the reference's algorithms in plain fp32 torch, each function citing the
reference file:line it follows, and are pinned bit-exactly against the unmodified reference Python
(tests/test_oracle_vs_reference.py, tests/golden/).  Only tests/, __graft_entry__.smoke() and bench.py's CPU-baseline /
--impl reference arm may import this package; the product (show-o_b200/) never does.
"""
