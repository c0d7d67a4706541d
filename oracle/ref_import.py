"""Import helpers for running the REFERENCE ITSELF in the build container (test infrastructure; the GPU box has no reference
checkout).  The reference's evaluation modules import plotting / physics packages that are not installed here (matplotlib,
energyflow, coffea, awkward, jetnet, mplhep); none of them is used by the arithmetic this repo pins, so they are replaced by
inert stub modules before the import."""
from __future__ import annotations

import sys
import types
import warnings

STUBS = ("matplotlib", "matplotlib.pyplot", "matplotlib.colors", "matplotlib.cm", "matplotlib.ticker", "matplotlib.patches",
         "energyflow", "coffea", "coffea.nanoevents", "coffea.nanoevents.methods", "awkward", "jetnet", "jetnet.datasets",
         "jetnet.losses", "mplhep")


class _Stub(types.ModuleType):
    __path__ = []

    def __getattr__(self, k):
        if k.startswith("__"):
            raise AttributeError(k)
        m = _Stub(self.__name__ + "." + k)
        sys.modules[m.__name__] = m
        setattr(self, k, m)
        return m

    def __call__(self, *a, **k):
        return self


def add_reference_to_path(ref="/root/reference"):
    warnings.filterwarnings("ignore")
    for name in STUBS:
        sys.modules.setdefault(name, _Stub(name))
    if ref not in sys.path:
        sys.path.insert(0, ref)
