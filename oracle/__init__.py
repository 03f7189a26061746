"""CPU oracle for the GPAD hot path -- TEST INFRASTRUCTURE ONLY.

Only tests/, __graft_entry__.smoke() and bench.py's cpu_baseline / --impl reference legs may
import this package; the product (gpu-dualgradient-mpc_b200/) never does.  See gpad_oracle.h.
"""
from .binding import (Oracle, RefCuda, RefLib, build, have_ref, have_refcuda, schedule, STATUS_NAMES)  # noqa: F401
