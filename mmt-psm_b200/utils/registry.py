"""String-keyed plug-in registry — mirrors `maskrcnn_benchmark.utils.registry.Registry`
(reference utils/registry.py:9-45): `REG.register("name", obj)` or `@REG.register("name")`."""


class Registry(dict):
    def register(self, name, obj=None):
        if obj is not None:
            self._add(name, obj)
            return obj

        def deco(fn):
            self._add(name, fn)
            return fn
        return deco

    def _add(self, name, obj):
        if name in self:
            raise KeyError("%r already registered" % name)
        self[name] = obj
