"""CPU oracle for the ReGAT implicit-relation hot path.

TEST INFRASTRUCTURE ONLY.  Nothing in the product package (tf_vqa_regat_b200/)
may import this package; only tests/, __graft_entry__.smoke() and bench.py's
cpu_baseline / --impl reference legs do, and there only as the checker or the
reported CPU baseline.

Pinning status:
  * stage 1 (position_emb.py)  -- PINNED: checked against outputs of the
    reference's own /root/reference/model/position_emb.py, committed as
    tests/golden/stage1_*.npz by oracle/make_golden.py.
  * stages 2-3, loss, optimizer -- PARITY UNPINNED: the reference needs
    TensorFlow, which is not installable here, and ships no tests or golden
    vectors.  The restatement follows the reference source op for op and is
    cross-checked two ways (NumPy vs torch-CPU autograd), nothing more.
"""
