"""Drop-in alias of the reference module name: ``from simple_models import *``."""
from federated_pytorch_test_b200.models import *  # noqa: F401,F403
from federated_pytorch_test_b200.models import __all__  # noqa: F401
