"""Minimal ``nodes`` package so that ``from nodes.V_nodes.v5_texture_ela import run`` resolves exactly as in the
reference tree (nodes/__init__.py there also star-imports every other node; those are out of scope here).

When the replacement module is dropped into the reference checkout the reference's own
``nodes/__init__.py`` is used instead and this file is not needed.
"""
import json
import os


def dump_node_debug(state: dict, node_name: str, payload: dict) -> None:
    """Same contract as the reference helper (nodes/__init__.py:5-22): ``<node>_debug.json`` + a line in
    ``debug_log.txt`` under ``state['data_dir']``; never raises."""
    data_dir = state.get("data_dir")
    if not data_dir:
        return
    try:
        with open(os.path.join(data_dir, f"{node_name}_debug.json"), "w") as f:
            json.dump(payload, f, indent=2)
        with open(os.path.join(data_dir, "debug_log.txt"), "a") as f:
            f.write(f"Node {node_name} completed. Keys: {list(payload.keys())}\n")
    except Exception:
        pass
