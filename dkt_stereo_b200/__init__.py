"""B200-native inference engine for the DKT-Stereo hot path.

``__models__`` mirrors the reference registry (meta_arch/__init__.py:7-12) for the two model families
this engine serves; unlike the reference it registers ``IGEVStereo`` too (the reference's configs
name it but its registry forgets it)."""
__version__ = "0.2.0"


class _LazyModels(dict):
    """name -> class, imported on first use so that `import dkt_stereo_b200` stays light."""

    _TABLE = {"RAFTStereo": ("raft_stereo", "RAFTStereo"), "IGEVStereo": ("igev_stereo", "IGEVStereo")}

    def __missing__(self, key):
        if key not in self._TABLE:
            raise KeyError(f"{key!r} is not served by the B200 engine (available: {sorted(self._TABLE)})")
        import importlib
        mod, cls = self._TABLE[key]
        self[key] = getattr(importlib.import_module(f"{__name__}.{mod}"), cls)
        return self[key]

    def __contains__(self, key):
        return key in self._TABLE

    def names(self):
        return sorted(self._TABLE)


__models__ = _LazyModels()
