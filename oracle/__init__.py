"""TEST INFRASTRUCTURE ONLY: CPU restatement of the reference decoder (see ldpc_oracle.h).

Only tests/, __graft_entry__.smoke() and bench.py's cpu_baseline / --impl reference legs
may import this package.  The product package (ldpc_decoder_b200) never does.
"""
