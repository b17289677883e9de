from . import Extension, ExtensionManager  # noqa: F401
