"""oracle/ — TEST INFRASTRUCTURE ONLY.

CPU restatement (plain PyTorch functional ops, fp32 or fp64) of the reference algorithm on the hot path.
Only tests/, __graft_entry__.smoke() and bench.py's cpu_baseline / --impl reference legs may import this package;
the product package (segmentation-and-classification-of-brain-tumor-using-3d-unet_b200/) never does.

Parity status: PINNED — every function here is checked in tests/test_oracle_golden.py against golden vectors generated
by executing the reference's own classes (sliced out of /root/reference/main.py, losses.py, training.py by
tests/golden/make_golden.py, run in the build container where the reference is mounted).
"""
