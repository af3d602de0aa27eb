"""loader of the pybind11 harness over the C++ host layer"""
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def load():
    from cozo_b200 import build
    build.build()
    build.build_host()
    p = os.path.join(ROOT, "cozo_b200", "host")
    if p not in sys.path:
        sys.path.insert(0, p)
    import _cozo_host
    return _cozo_host
