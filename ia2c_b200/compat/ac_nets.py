"""Drop-in for the reference's ``ac_nets`` module.  The scripts do ``from ac_nets import *`` and rely on
it to provide ``torch``, ``nn``, ``F``, ``np``, ``Adam``, ``Categorical`` and ``hidden_size`` as well
(ia2c.py:22, a2c_org_test.py:19), so they are re-exported here."""
import numpy as np  # noqa: F401
import torch  # noqa: F401
import torch.nn as nn  # noqa: F401
import torch.nn.functional as F  # noqa: F401
from torch.distributions import Categorical  # noqa: F401

from ia2c_b200.nets import ActorNetwork, Adam, CriticNetwork, NeuralNet, hidden_size  # noqa: F401
