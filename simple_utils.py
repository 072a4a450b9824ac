"""Drop-in alias of the reference module name: ``from simple_utils import *``."""
from federated_pytorch_test_b200.utils.simple_utils import *  # noqa: F401,F403
