"""Import shim (test infrastructure): `unpackable.unpack` is the only symbol the reference's hot path
uses (agents/dreamer_v2.py:9, utils/replay_buffer.py:4)."""
import dataclasses


def unpack(obj):
    return tuple(getattr(obj, f.name) for f in dataclasses.fields(obj))
