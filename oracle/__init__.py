"""oracle/ -- TEST INFRASTRUCTURE (CPU restatement of the reference's block kernels, fixture
generators, reference import helpers).  Only tests/, __graft_entry__.smoke() and bench.py's
cpu_baseline / --impl reference legs may import from here; the product never does."""
