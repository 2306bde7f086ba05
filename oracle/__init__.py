"""CPU oracle (test infrastructure only).  See oracle/snp_oracle.c for the header and the pin.

Only tests/, __graft_entry__.smoke() and bench.py's cpu_baseline / --impl reference legs import this.
"""
from .oracle import (OracleConfig, update_humans, imitation_steps, sim_update_steps, checks, laser, lookahead, robot_push_out, SFMS, type_code, default_params,  # noqa: F401
                     N_STATE, N_PARAMS)
