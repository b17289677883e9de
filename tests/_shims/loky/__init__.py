"""Minimal stand-in for `loky` (test infrastructure only)."""
import multiprocessing
from concurrent.futures import ProcessPoolExecutor


def cpu_count():
    return multiprocessing.cpu_count()


def get_reusable_executor(max_workers=None, **kw):
    return ProcessPoolExecutor(max_workers=max_workers)


class backend:
    @staticmethod
    def get_context():
        return multiprocessing.get_context()
