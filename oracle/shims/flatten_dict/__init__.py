"""Import shim (test infrastructure) for utils/logger.py."""


def flatten(d, reducer=None):
    return d
