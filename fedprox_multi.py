#!/usr/bin/env python
"""`python fedprox_multi.py [--K 8 --use_resnet ...]` — same script name as the reference; see
federated_pytorch_test_b200/api/fedprox_multi.py for the implementation and knob list."""
from federated_pytorch_test_b200.api.fedprox_multi import main

if __name__ == "__main__":
    main()
