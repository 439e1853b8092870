"""Drop-in for the reference's ``Org`` module (``from Org import Org``; gym entry point ``"Org:Org"``).

Put ``ia2c_b200/compat`` on ``sys.path`` (before the reference's directory) and ``a2c_org_test.py`` /
``ia2c.py`` import this instead of Org.py.  Module-level flags mirror Org.py:6-10; only the shipped
default (MEM=True, MEM_SIZE=1, STATE_VISIBLE=False) is implemented.
"""
from ia2c_b200.org_env import Org, OrgVecEnv  # noqa: F401

MEM_SIZE = 1
STATE_VISIBLE = False
MEM = True
STATE = False
LSTM = False
