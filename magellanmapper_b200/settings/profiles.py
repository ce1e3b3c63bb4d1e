"""Layered settings dictionaries (mirror of ``magmap/settings/profiles.py``).

A ``SettingsDict`` starts from defaults and is modified by named "modifier"
groups or YAML files applied in order, later ones winning
(``profiles.py:218-240``).
"""
from __future__ import annotations

import os
from typing import Dict, Optional, Sequence

import yaml


class SettingsDict(dict):
    PATH_PROFILES = "profiles"
    NAME_KEY = "settings_name"
    DEFAULT_NAME = "default"
    _YAML_EXT = (".yml", ".yaml")

    def __init__(self, *args, **kwargs):
        super().__init__()
        self[self.NAME_KEY] = self.DEFAULT_NAME
        self.profiles: Dict[str, Dict] = {}
        self.timestamps: Dict[str, float] = {}
        self.delimiter = ","
        self.update(*args, **kwargs)

    # -- modifiers ---------------------------------------------------------
    def modify_settings(self, mods: Dict) -> None:
        """Overlay ``mods``: nested dicts are merged, everything else replaced;
        unknown keys fall back to attributes (profiles.py:122-158)."""
        for key, val in mods.items():
            if key in self:
                if isinstance(self[key], dict) and isinstance(val, dict):
                    self[key].update(val)
                else:
                    self[key] = val
            elif hasattr(self, key):
                cur = getattr(self, key)
                if isinstance(cur, dict) and isinstance(val, dict):
                    cur.update(val)
                else:
                    setattr(self, key, val)
            else:
                # the reference logs and ignores unknown keys; new keys from a
                # YAML file are still useful to downstream readers, so keep them
                self[key] = val

    def get_profile(self, profile_name: str) -> Optional[Dict]:
        """Resolve a modifier by name or YAML path (profiles.py:160-216)."""
        if os.path.splitext(profile_name)[1].lower() in self._YAML_EXT:
            path = os.path.join(self.PATH_PROFILES, profile_name)
            if not os.path.exists(path):
                path = profile_name
            if not os.path.exists(path):
                # profiles shipped with this package
                path = os.path.join(os.path.dirname(os.path.dirname(__file__)),
                                    "profiles", os.path.basename(profile_name))
                if not os.path.exists(path):
                    return None
            self.timestamps[path] = os.path.getmtime(path)
            mods: Dict = {}
            with open(path) as f:
                for doc in yaml.safe_load_all(f):
                    if doc:
                        mods.update(doc)
            return mods
        if profile_name == self.DEFAULT_NAME:
            return self.__class__()
        return self.profiles.get(profile_name)

    def add_profiles(self, names_str: str) -> None:
        for name in names_str.split(self.delimiter):
            mods = self.get_profile(name)
            if mods:
                self[self.NAME_KEY] += self.delimiter + name
                self.modify_settings(mods)

    def check_file_changed(self) -> bool:
        return any(t < os.path.getmtime(p) for p, t in self.timestamps.items())

    def refresh_profile(self, check_timestamp: bool = False) -> None:
        if not check_timestamp or self.check_file_changed():
            names = self[self.NAME_KEY]
            self.__init__()
            self.add_profiles(names)

    @staticmethod
    def is_identical_settings(profs: Sequence["SettingsDict"], keys: Sequence[str]) -> bool:
        """True when every profile agrees with the first on ``keys``
        (profiles.py:272-297)."""
        first = None
        for prof in profs:
            if first is None:
                first = prof
                continue
            if any(first[k] != prof[k] for k in keys):
                return False
        return True
