"""CPU oracle for the MHAP sketch + overlap-search path.  TEST INFRASTRUCTURE ONLY.

Only tests/, __graft_entry__.smoke() and bench.py's cpu_baseline / --impl reference legs may
import this package.  Parity is UNPINNED by the reference's own tests (it ships none and no JVM
exists here); see oracle/mhap_oracle.h.
"""
