#!/usr/bin/env python
"""`python federated_vae.py [--K 8 --use_resnet ...]` — same script name as the reference; see
federated_pytorch_test_b200/api/federated_vae.py for the implementation and knob list."""
from federated_pytorch_test_b200.api.federated_vae import main

if __name__ == "__main__":
    main()
