"""Drop-in alias of the reference module name: ``from lbfgsnew import LBFGSNew``."""
from federated_pytorch_test_b200.optim.lbfgsnew import LBFGSNew  # noqa: F401
