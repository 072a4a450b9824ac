#!/usr/bin/env python
"""`python no_consensus_multi.py [--K 8 --use_resnet ...]` — same script name as the reference; see
federated_pytorch_test_b200/api/no_consensus_multi.py for the implementation and knob list."""
from federated_pytorch_test_b200.api.no_consensus_multi import main

if __name__ == "__main__":
    main()
