"""``jaxmarl.wrappers.baselines.JaxMARLWrapper`` stand-in: stores ``_env`` and forwards attributes."""


class JaxMARLWrapper:
    def __init__(self, env):
        self._env = env

    def __getattr__(self, name):
        return getattr(self._env, name)
