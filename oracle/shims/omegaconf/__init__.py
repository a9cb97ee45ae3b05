"""Import shim (test infrastructure): utils/logger.py:4-5 imports these names at module load."""


class DictConfig(dict):
    pass


class OmegaConf:
    to_container = staticmethod(lambda c, **k: dict(c))
