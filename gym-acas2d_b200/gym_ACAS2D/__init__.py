"""B200-native drop-in for the ``gym_ACAS2D`` package of Christos-14/gym-ACAS2D.

Importing the package registers ``ACAS2D-v0`` exactly like the reference
(``gym_ACAS2D/__init__.py:3-6``) when ``gym`` (or ``gymnasium``) is installed; without either,
``gym_ACAS2D.make("ACAS2D-v0")`` resolves the same id through a local registry.
"""
ENV_ID = "ACAS2D-v0"
ENTRY_POINT = "gym_ACAS2D.envs:ACAS2DEnv"

_registry = {ENV_ID: ENTRY_POINT}
REGISTERED_WITH = None

try:                                            # pragma: no cover - not installed in the build image
    from gym.envs.registration import register
    register(id=ENV_ID, entry_point=ENTRY_POINT)
    REGISTERED_WITH = "gym"
except Exception:                               # noqa: BLE001
    try:
        from gymnasium.envs.registration import register
        register(id=ENV_ID, entry_point=ENTRY_POINT, disable_env_checker=True)
        REGISTERED_WITH = "gymnasium"
    except Exception:                           # noqa: BLE001
        pass


def make(env_id: str = ENV_ID, **kwargs):
    """``gym.make`` stand-in for machines without gym: instantiate a registered id."""
    import importlib
    if env_id not in _registry:
        raise KeyError(f"unknown environment id {env_id!r}")
    module, _, cls = _registry[env_id].partition(":")
    return getattr(importlib.import_module(module), cls)(**kwargs)
