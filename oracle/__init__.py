"""CPU oracle for the cosdata distance hot path -- TEST INFRASTRUCTURE ONLY.

Only tests/, __graft_entry__.smoke() and bench.py's cpu_baseline /
``--impl reference`` legs may import this package.  The product package
(cosdata_b200) never does.  See oracle/cosdata_oracle.h for the parity status.
"""
from .pyoracle import *  # noqa: F401,F403
