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


def _route_make_vec():
    """With the REAL gymnasium installed, ``gym.make_vec("Org-v0", num_envs=E)`` (ia2c.py:42) would spawn E worker
    processes around single-env instances.  Route ids whose entry point is this drop-in ``Org`` to the batched GPU
    env instead; every other id keeps gymnasium's behaviour."""
    try:
        import gymnasium as gym
    except Exception:
        return
    if getattr(gym, "__version__", "").endswith("ia2c_b200.shim") or getattr(gym.make_vec, "_ia2c_routed", False):
        return
    original = gym.make_vec

    def make_vec(id, num_envs=1, *args, **kwargs):
        try:
            spec = gym.spec(id) if isinstance(id, str) else id
            entry = spec.entry_point
            if entry == "Org:Org" or entry is Org:
                return OrgVecEnv(num_envs, n_agents=2, max_episode_steps=spec.max_episode_steps)
        except Exception:
            pass
        return original(id, num_envs, *args, **kwargs)

    make_vec._ia2c_routed = True
    gym.make_vec = make_vec


_route_make_vec()
