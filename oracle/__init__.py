"""ORACLE package: CPU restatement of the reference algorithms (test infrastructure only).

Only tests/, __graft_entry__.smoke() and bench.py's cpu_baseline / --impl reference legs may import this.
The product package (era_boojum_b200) never does.
"""
from .oracle import *  # noqa: F401,F403
