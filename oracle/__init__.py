"""CPU oracle for the conv-net train-step hot path.  TEST INFRASTRUCTURE ONLY.

`oracle.port` binds the plain-C restatement (oracle/cnn_oracle.c); `oracle.ref`
binds oracle/_ref/libcnn_ref.so, the reference's own sources compiled in place.
Only tests/, bench.py's cpu_baseline / --impl reference leg and
__graft_entry__.smoke() may import this package; cnn_b200/ never does.
"""
