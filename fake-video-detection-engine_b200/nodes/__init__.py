"""Minimal ``nodes`` package so that ``from nodes.V_nodes.v5_texture_ela import run`` resolves as in the reference tree.

Only the one helper V5 needs lives here; the reference's own ``nodes/__init__.py`` additionally star-imports every other
node (out of scope). When the replacement module is dropped into a reference checkout the reference's
package is used and this file is irrelevant.
"""
import json
from pathlib import Path


def dump_node_debug(state: dict, node_name: str, payload: dict) -> None:
    """Contract of the reference helper (nodes/__init__.py:5-22): best-effort ``<node>_debug.json`` plus one line appended
    to ``debug_log.txt`` under ``state['data_dir']``; silently does nothing without a data_dir or on any I/O error."""
    root = state.get("data_dir")
    if not root:
        return
    try:
        base = Path(root)
        (base / f"{node_name}_debug.json").write_text(json.dumps(payload, indent=2))
        with (base / "debug_log.txt").open("a") as log:
            log.write(f"Node {node_name} completed. Keys: {list(payload.keys())}\n")
    except Exception:
        return
