"""Runs the reference's UNMODIFIED tests/test_v5_texture_ela.py against a reference checkout whose V5 node module has been replaced
by this repo's. Test infrastructure, used by tests/test_reference_dropin.py in a subprocess.

    python tests/dropin_runner.py <tree> [unittest test names ...]

<tree> is a copy of the reference's ``nodes/`` and ``tests/test_v5_texture_ela.py`` with nodes/V_nodes/v5_texture_ela.py overwritten.
The reference's package __init__ files star-import every node, and most of their third-party dependencies (whisper, easyocr,
face_alignment, ...) are not installed in this image: every module that cannot be found is replaced by an empty stand-in before the
import is retried — the same thing oracle/ref_loader.py does for the golden generator, without touching a line of the tree."""
import importlib
import os
import sys
import types
import unittest
from unittest.mock import MagicMock


class _Stub(types.ModuleType):
    __path__ = []                                         # lets `import a.b` resolve through the finder below

    def __getattr__(self, name):
        if name.startswith("__"):
            raise AttributeError(name)
        return MagicMock(name=f"{self.__name__}.{name}")


def main():
    tree = os.path.abspath(sys.argv[1])
    names = sys.argv[2:]
    repo = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
    sys.path[:0] = [tree, os.path.join(repo, "fake-video-detection-engine_b200")]   # the tree's `nodes` first, then `v5ela`
    sys.dont_write_bytecode = True
    stubbed = []
    for _ in range(64):
        try:
            mod = importlib.import_module("nodes.V_nodes.v5_texture_ela")
            break
        except ModuleNotFoundError as e:
            if not e.name or e.name.startswith("nodes") or e.name.startswith("v5ela"):
                raise
            for k in [k for k in sys.modules if k == "nodes" or k.startswith("nodes.")]:
                del sys.modules[k]
            sys.modules[e.name] = _Stub(e.name)
            stubbed.append(e.name)
    else:
        raise SystemExit("could not import the node module")
    assert os.path.abspath(mod.__file__).startswith(tree), mod.__file__
    assert "v5ela" in open(mod.__file__).read(), "the tree still holds the reference's own node module"
    print("dropin: stand-ins for", sorted(stubbed))
    test_mod = importlib.import_module("tests.test_v5_texture_ela") if os.path.exists(os.path.join(tree, "tests", "__init__.py")) else None
    if test_mod is None:
        spec = importlib.util.spec_from_file_location("test_v5_texture_ela", os.path.join(tree, "tests", "test_v5_texture_ela.py"))
        test_mod = importlib.util.module_from_spec(spec)
        sys.modules["test_v5_texture_ela"] = test_mod
        spec.loader.exec_module(test_mod)
    loader = unittest.defaultTestLoader
    suite = loader.loadTestsFromNames(names, test_mod) if names else loader.loadTestsFromModule(test_mod)
    result = unittest.TextTestRunner(verbosity=2).run(suite)
    print(f"dropin: ran {result.testsRun} failures {len(result.failures)} errors {len(result.errors)}")
    sys.exit(0 if result.wasSuccessful() and result.testsRun > 0 else 1)


if __name__ == "__main__":
    main()
