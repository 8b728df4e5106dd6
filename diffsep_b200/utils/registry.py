"""Name -> class plugin registry: the extension point of the predictor / corrector / SDE families
(same behaviour as the reference ``utils/registry.py:5-36``: double registration warns and
replaces, unknown names raise ``ValueError``)."""
from __future__ import annotations

import warnings


class Registry:
    def __init__(self, managed_thing: str):
        self.managed_thing = managed_thing
        self._registry = {}

    def register(self, name: str):
        def deco(cls):
            if name in self._registry:
                warnings.warn(f"{self.managed_thing} with name '{name}' doubly registered, "
                              "old class will be replaced.")
            self._registry[name] = cls
            return cls
        return deco

    def get_by_name(self, name: str):
        try:
            return self._registry[name]
        except KeyError:
            raise ValueError(f"{self.managed_thing} with name '{name}' unknown.") from None

    def get_all_names(self):
        return list(self._registry.keys())
