"""Import the UNMODIFIED reference V5 node. TEST INFRASTRUCTURE ONLY (see oracle/__init__.py).

``import nodes`` fails in this image because ``nodes/__init__.py:25-29`` star-imports every node package and most
of their third-party dependencies (moviepy, whisper, easyocr, ...) are not installed. V5's own imports are all
present, so we register a stub ``nodes`` package that carries only ``dump_node_debug`` (executed from the head of
the reference's ``nodes/__init__.py``, lines 1-22) and then load ``nodes/V_nodes/v5_texture_ela.py`` from where it
lies. Nothing is copied; ``/root/reference`` only exists in the build container, so callers must handle
``ReferenceUnavailable`` (the golden vectors under tests/golden/ are what travels to the GPU box).
"""
from __future__ import annotations

import importlib.util
import os
import sys
import types

REF_ROOT = os.environ.get("V5ELA_REFERENCE_ROOT", "/root/reference")


class ReferenceUnavailable(RuntimeError):
    pass


def load_reference_v5(ref_root: str = REF_ROOT, module_name: str = "_ref_nodes"):
    """Return the reference's ``v5_texture_ela`` module object (its ``run`` is the node).

    The stub package is registered under a private top-level name so it cannot collide with the replacement
    ``nodes`` package of this repo; ``from nodes import dump_node_debug`` inside the reference module is served
    by temporarily aliasing ``sys.modules['nodes']`` while it executes.
    """
    node_py = os.path.join(ref_root, "nodes", "V_nodes", "v5_texture_ela.py")
    init_py = os.path.join(ref_root, "nodes", "__init__.py")
    if not (os.path.exists(node_py) and os.path.exists(init_py)):
        raise ReferenceUnavailable(f"reference not present at {ref_root}")
    sys.dont_write_bytecode = True
    with open(init_py) as f:
        head = "".join(f.readlines()[:22])          # dump_node_debug only; stops before the star-imports
    stub = types.ModuleType(module_name)
    stub.__path__ = [os.path.join(ref_root, "nodes")]
    exec(compile(head, init_py, "exec"), stub.__dict__)
    vpkg = types.ModuleType(module_name + ".V_nodes")
    vpkg.__path__ = [os.path.join(ref_root, "nodes", "V_nodes")]
    spec = importlib.util.spec_from_file_location(module_name + ".V_nodes.v5_texture_ela", node_py)
    mod = importlib.util.module_from_spec(spec)
    saved = sys.modules.get("nodes")
    sys.modules["nodes"] = stub
    try:
        spec.loader.exec_module(mod)
    finally:
        if saved is not None:
            sys.modules["nodes"] = saved
        else:
            sys.modules.pop("nodes", None)
    return mod
