from .common import *  # noqa: F401,F403  (the reference re-exports common from the package, dreamer/__init__.py:1)
