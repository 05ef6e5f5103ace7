"""Test infrastructure: the CPU restatement (oracle.py) and the reference's own CUDA
kernels behind a C ABI (ref_gpu.py).  Only tests/, __graft_entry__.smoke() and bench.py's
cpu_baseline / --impl reference legs may import this package; the product
(radfoam_b200/) never does."""
