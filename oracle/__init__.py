"""TEST INFRASTRUCTURE ONLY: CPU oracle for the CUDA path (see oracle/xq_oracle.h).

Only tests/, __graft_entry__.smoke() and bench.py's cpu_baseline / --impl reference legs may
import this package.  The product package cn_chess_ai_b200 never does.
"""
