"""TEST INFRASTRUCTURE ONLY — loads the *unmodified* reference (ExTrack 1.6.3) in-process.

Only ``tests/``, ``tests/golden/make_golden.py`` and ad-hoc validation scripts may import
this module.  It works only where ``/root/reference`` exists (the build container); the
GPU box has no reference tree, so nothing on the ``-m gpu`` / ``smoke()`` / ``bench.py``
paths may depend on it (they use the committed fixtures under ``tests/golden``).

The reference imports ``lmfit`` at module import (``extrack/tracking.py:31``) and lmfit is
absent from this image, so the repo's stand-in (``extrack_b200._lmfit_compat``) is
installed under the name ``lmfit`` first.  Only ``Parameters`` / ``.value`` are touched
by the likelihood path.
"""
from __future__ import annotations

import importlib.util
import os
import sys
import types

REFERENCE_ROOT = os.environ.get("EXTRACK_REFERENCE_ROOT", "/root/reference")


def reference_available() -> bool:
    return os.path.isfile(os.path.join(REFERENCE_ROOT, "extrack", "tracking.py"))


def _install_lmfit_stub():
    if "lmfit" in sys.modules:
        return
    try:
        import lmfit  # noqa: F401

        return
    except Exception:
        pass
    here = os.path.dirname(os.path.abspath(__file__))
    root = os.path.dirname(here)
    if root not in sys.path:
        sys.path.insert(0, root)
    from extrack_b200 import _lmfit_compat as compat

    stub = types.ModuleType("lmfit")
    stub.Parameters = compat.Parameters
    stub.Parameter = compat.Parameter
    stub.minimize = compat.minimize
    sys.modules["lmfit"] = stub


def _load(name: str, relpath: str):
    modname = "_extrack_reference_" + name
    if modname in sys.modules:
        return sys.modules[modname]
    _install_lmfit_stub()
    path = os.path.join(REFERENCE_ROOT, relpath)
    spec = importlib.util.spec_from_file_location(modname, path)
    mod = importlib.util.module_from_spec(spec)
    sys.modules[modname] = mod
    spec.loader.exec_module(mod)
    return mod


def load_tracking():
    """The reference's ``extrack/tracking.py`` as a module object (unmodified)."""
    return _load("tracking", "extrack/tracking.py")


def load_simulate():
    return _load("simulate_tracks", "extrack/simulate_tracks.py")


def load_readers():
    # readers.py imports xmltodict at module import (absent from this image): a stand-in with xmltodict's documented
    # mapping for attribute-only documents (attributes -> '@name' keys, repeated child elements -> a list, a single child
    # -> a dict), which is all read_trackmate_xml touches
    if "xmltodict" not in sys.modules:
        try:
            import xmltodict  # noqa: F401
        except Exception:
            stub = types.ModuleType("xmltodict")

            def parse(text, encoding="utf-8", **_kw):
                import xml.etree.ElementTree as ET

                def conv(el):
                    d = {"@" + k: v for k, v in el.attrib.items()}
                    for ch in el:
                        v = conv(ch)
                        if ch.tag in d:
                            if not isinstance(d[ch.tag], list):
                                d[ch.tag] = [d[ch.tag]]
                            d[ch.tag].append(v)
                        else:
                            d[ch.tag] = v
                    if el.text and el.text.strip():
                        if d:
                            d["#text"] = el.text.strip()
                        else:
                            return el.text.strip()
                    return d if d else None

                root = ET.fromstring(text.encode(encoding) if isinstance(text, str) else text)
                return {root.tag: conv(root)}

            stub.parse = parse
            sys.modules["xmltodict"] = stub
    return _load("readers", "extrack/readers.py")


def load_histograms():
    """``extrack/histograms.py`` (it imports ``extrack.tracking``: a package stub points at the loaded module)."""
    trk = load_tracking()
    if "extrack" not in sys.modules:
        pkg = types.ModuleType("extrack")
        pkg.__path__ = []  # a package, so that ``from extrack.tracking import ...`` resolves through sys.modules
        pkg.tracking = trk
        sys.modules["extrack"] = pkg
        sys.modules["extrack.tracking"] = trk
    return _load("histograms", "extrack/histograms.py")


def load_refined_localization():
    """``extrack/refined_localization.py`` (position refinement).  It imports ``extrack.tracking_0``, ``extrack.tracking``,
    ``extrack.exporters`` and, unconditionally, ``matplotlib.backends.backend_agg`` (absent from this image): package
    stubs point at the loaded reference modules and an empty stand-in satisfies the plotting import, which the
    refinement itself never touches."""
    trk = load_tracking()
    trk0 = _load("tracking_0", "extrack/tracking_0.py")
    exporters = _load("exporters", "extrack/exporters.py")
    if "extrack" not in sys.modules:
        pkg = types.ModuleType("extrack")
        pkg.__path__ = []
        sys.modules["extrack"] = pkg
    pkg = sys.modules["extrack"]
    for name, mod in (("tracking", trk), ("tracking_0", trk0), ("exporters", exporters)):
        setattr(pkg, name, mod)
        sys.modules["extrack." + name] = mod
    try:
        import matplotlib.backends.backend_agg  # noqa: F401
    except Exception:
        mpl = sys.modules.setdefault("matplotlib", types.ModuleType("matplotlib"))
        be = sys.modules.setdefault("matplotlib.backends", types.ModuleType("matplotlib.backends"))
        agg = types.ModuleType("matplotlib.backends.backend_agg")
        agg.FigureCanvasAgg = object
        sys.modules["matplotlib.backends.backend_agg"] = agg
        mpl.backends = be
        be.backend_agg = agg
    if not hasattr(__import__("numpy"), "product"):
        __import__("numpy").product = __import__("numpy").prod  # tracking_0.fuse_tracks_general predates numpy 2
    return _load("refined_localization", "extrack/refined_localization.py")
