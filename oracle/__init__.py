"""ORACLE -- test infrastructure only.

CPU restatement of the reference's algorithms for the hot path (BN254 field/curve/MSM, multilinear
polynomials, sumcheck round loops, Keccak transcript):
    bn254.py, poly.py, sumcheck.py, transcript.py   Python big-integer truth
    cpu_ref.cpp (+ cpu_ref.py)                      C++17 restatement "as written": fast checker and
                                                    the timed CPU baseline
Only tests/, __graft_entry__.smoke() and bench.py's cpu_baseline / --impl reference legs may import
this package.  Parity status: PINNED against the reference's own unit-test vectors and logs/zolt.log
KATs (tests/golden/, tests/test_oracle_golden.py); MSM values are pinned mathematically only (the
reference's tests discard MSM results).
"""
