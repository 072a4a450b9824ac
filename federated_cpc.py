#!/usr/bin/env python
"""`python federated_cpc.py [--K 8 --use_resnet ...]` — same script name as the reference; see
federated_pytorch_test_b200/api/federated_cpc.py for the implementation and knob list."""
from federated_pytorch_test_b200.api.federated_cpc import main

if __name__ == "__main__":
    main()
