"""TEST INFRASTRUCTURE -- the CPU oracle.  Only tests/, __graft_entry__.smoke() and bench.py's
cpu_baseline / --impl reference legs may import this package; the product never does."""
from .oracle import *  # noqa: F401,F403
