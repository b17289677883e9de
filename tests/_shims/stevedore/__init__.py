"""Minimal stand-in for `stevedore` so the reference imports in the build container.

Entry-point tables are read from the reference's own pyproject.toml (the
reference is not installed, so importlib.metadata knows nothing about it).
Test infrastructure only.
"""
import importlib
import os
import tomllib

_REF = os.environ.get("COGENT3_REFERENCE_ROOT", "/root/reference")


def _entry_points(namespace):
    table = None
    # <root>/pyproject.toml (baseline/_ref) or <root>/../pyproject.toml (root = /root/reference/src)
    for path in (os.path.join(_REF, "pyproject.toml"), os.path.join(os.path.dirname(_REF.rstrip("/")), "pyproject.toml")):
        try:
            with open(path, "rb") as f:
                table = tomllib.load(f)
            break
        except OSError:
            continue
    if table is None:
        return {}
    return table.get("project", {}).get("entry-points", {}).get(namespace, {})


class Extension:
    def __init__(self, name, target):
        self.name = name
        self.module_name, self.attr = target.split(":")
        self.entry_point = target
        self._plugin = None
        self.obj = None

    @property
    def plugin(self):
        if self._plugin is None:
            mod = importlib.import_module(self.module_name)
            self._plugin = getattr(mod, self.attr)
        return self._plugin


class ExtensionManager:
    def __init__(self, namespace, invoke_on_load=False, **kw):
        self.namespace = namespace
        self.extensions = [Extension(k, v) for k, v in _entry_points(namespace).items()]

    def __iter__(self):
        return iter(self.extensions)

    def __getitem__(self, name):
        for e in self.extensions:
            if e.name == name:
                return e
        raise KeyError(name)

    def names(self):
        return [e.name for e in self.extensions]

    def __contains__(self, name):
        return name in self.names()


class _exception:
    class NoMatches(RuntimeError):
        pass


exception = _exception


class _hook:
    class HookManager(ExtensionManager):
        def __init__(self, namespace, name, invoke_on_load=False, **kw):
            super().__init__(namespace)
            self.extensions = [e for e in self.extensions if e.name == name]


hook = _hook


class _driver:
    class DriverManager(ExtensionManager):
        def __init__(self, namespace, name, invoke_on_load=False, **kw):
            super().__init__(namespace)
            self.extensions = [e for e in self.extensions if e.name == name]
            if not self.extensions:
                raise exception.NoMatches(f"No {namespace!r} driver found, looking for {name!r}")

        @property
        def driver(self):
            return self.extensions[0].plugin


driver = _driver
