"""oracle — CPU checkers (TEST INFRASTRUCTURE ONLY; see oracle/README.md).

Only tests/, __graft_entry__.smoke() and bench.py's cpu_baseline / --impl reference legs may import this package.
The product (gslam_b200/) never does, and fails loudly when its CUDA library is missing.
"""
from .oracle import *  # noqa: F401,F403
