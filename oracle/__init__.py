"""oracle/ -- TEST INFRASTRUCTURE ONLY.

CPU restatement of the reference algorithm (mph_oracle.c -> liboracle.so) and the untouched
reference itself compiled in place (oracle/_ref, see build_ref.sh).  Only tests/,
__graft_entry__.smoke() and bench.py's cpu_baseline / --impl reference legs may import this package;
the product (particlemethod_fsi_b200/) never does.
"""
