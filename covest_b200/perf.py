"""Nested wall-clock timers reporting to stderr.

Only the message text is a contract with the reference ("<what> took <seconds> seconds.", and
"Function <name> (<module>) took ..." for decorated functions; covest/perf.py:52-68): tools that
scrape covest's stderr keep working.  The mechanism is this package's own: one `Timer` per region,
the nesting depth kept in a thread-local counter (the multi-start refinement runs regions in
threads), monotonic clock, and the message is emitted even when the region raises.
"""
import functools
import sys
import threading
import time

_depth = threading.local()


class Timer:
    """Context manager: prints '<label> took <seconds> seconds.' on exit, indented by nesting."""

    def __init__(self, label, out=None):
        self.label = label
        self.out = out
        self.seconds = None

    def __enter__(self):
        self.level = getattr(_depth, 'n', 0)
        _depth.n = self.level + 1
        self.t0 = time.perf_counter()
        return self

    def __exit__(self, *exc):
        self.seconds = time.perf_counter() - self.t0
        _depth.n = self.level
        prefix = '' if self.level == 0 else '| ' * (self.level - 1) + '+'
        (self.out or sys.stderr).write('%s%s took %s seconds.\n' % (prefix, self.label, self.seconds))
        return False


def running_time(text):
    """`with running_time('First optimization'): ...` (covest/covest.py:54, :65; grid.py:62)."""
    return Timer(text)


def running_time_decorator(fn):
    """Times every call of `fn` (covest/covest.py:99, grid.py:17)."""
    label = 'Function %s (%s)' % (fn.__name__, fn.__module__)

    @functools.wraps(fn)
    def timed(*args, **kwargs):
        with Timer(label):
            return fn(*args, **kwargs)
    return timed
