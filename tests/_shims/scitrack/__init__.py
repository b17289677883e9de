"""Minimal stand-in for `scitrack` (test infrastructure only)."""
import hashlib


class CachingLogger:
    def __init__(self, *a, **kw):
        self.log_file_path = None
        self._messages = []

    def log_message(self, msg, label=None):
        self._messages.append((label, msg))

    def log_args(self, *a, **kw):
        pass

    def log_versions(self, *a, **kw):
        pass

    def input_file(self, *a, **kw):
        pass

    def output_file(self, *a, **kw):
        pass

    def shutdown(self):
        pass


def get_text_hexdigest(data):
    if isinstance(data, str):
        data = data.encode("utf8")
    return hashlib.md5(data).hexdigest()


def get_file_hexdigest(path):
    with open(path, "rb") as f:
        return hashlib.md5(f.read()).hexdigest()
