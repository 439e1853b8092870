import importlib

_REGISTRY = {}


class Env:
    metadata = {}


def register(id, entry_point=None, max_episode_steps=None, **kwargs):
    _REGISTRY[id] = dict(entry_point=entry_point, max_episode_steps=max_episode_steps, kwargs=kwargs)


def _load(entry_point):
    if callable(entry_point):
        return entry_point
    mod, attr = entry_point.split(":")
    return getattr(importlib.import_module(mod), attr)


def make(id, **kwargs):
    return _load(_REGISTRY[id]["entry_point"])()


def make_vec(id, num_envs=1, **kwargs):
    from ia2c_b200.org_env import Org, OrgVecEnv

    spec = _REGISTRY[id]
    cls = _load(spec["entry_point"])
    if cls is Org:
        return OrgVecEnv(num_envs, n_agents=2, max_episode_steps=spec["max_episode_steps"])
    raise NotImplementedError(f"the gymnasium shim only vectorises the Org domain (got entry point {spec['entry_point']!r}); "
                              "install the real gymnasium for other environments")
