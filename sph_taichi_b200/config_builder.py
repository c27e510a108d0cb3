"""Scene configuration reader.

Mirrors the reference's ``SimConfig`` accessors (reference ``config_builder.py:4-37``):
a JSON scene with a ``Configuration`` table and optional ``FluidBlocks`` /
``RigidBlocks`` / ``RigidBodies`` lists.  In addition to a file path this class
accepts an already-parsed ``dict`` so that synthetic benchmark scenes
(``sph_taichi_b200.scene``) do not have to touch the filesystem.
"""
from __future__ import annotations

import copy
import json
import os


class SimConfig:
    def __init__(self, scene_file_path, verbose: bool = False) -> None:
        if isinstance(scene_file_path, dict):
            self.config = copy.deepcopy(scene_file_path)
            self.scene_dir = os.getcwd()
        else:
            with open(scene_file_path, "r") as fh:
                self.config = json.load(fh)
            self.scene_dir = os.path.dirname(os.path.abspath(scene_file_path))
        if "Configuration" not in self.config:
            raise KeyError("scene has no 'Configuration' table")
        if verbose:
            print(self.config)

    # -- reference API -----------------------------------------------------
    def get_cfg(self, name, enforce_exist: bool = False):
        table = self.config["Configuration"]
        if name in table:
            return table[name]
        if enforce_exist:
            # the reference asserts here (config_builder.py:12-17)
            raise AssertionError(f"required configuration key '{name}' is missing")
        return None

    def _section(self, key):
        return self.config.get(key, [])

    def get_rigid_bodies(self):
        return self._section("RigidBodies")

    def get_rigid_blocks(self):
        return self._section("RigidBlocks")

    def get_fluid_blocks(self):
        return self._section("FluidBlocks")
