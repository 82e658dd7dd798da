"""Wall-clock timers that report to stderr in the reference's format
("... took {time} seconds.", covest/perf.py)."""
import sys
import time
from contextlib import contextmanager

stack = []
messages = []


def indent():
    if len(stack) < 2:
        return ''
    return '| ' * (len(stack) - 2) + '+'


def push(cnt=1):
    for _ in range(cnt):
        stack.append(time.time())


def pop(cnt=1):
    last = None
    for _ in range(cnt):
        last = stack.pop()
    return last


def replace():
    pop()
    push()


def get_time(back=0):
    n = len(stack)
    if back >= n or back == -1:
        back = n - 1
    return stack[n - back - 1]


def print_all():
    global messages
    for m in messages:
        sys.stderr.write(m + '\n')
    messages = []


def msg(message, back=0):
    elapsed = time.time() - get_time(back)
    messages.append(indent() + message.format(time=elapsed))
    print_all()


def running_time_decorator(fn):
    def wrapped(*args, **kwargs):
        push(2)
        try:
            return fn(*args, **kwargs)
        finally:
            msg('Function ' + fn.__name__ + ' (' + fn.__module__ + ') took {time} seconds.', 1)
            pop(2)
    wrapped.__name__ = fn.__name__
    wrapped.__doc__ = fn.__doc__
    return wrapped


@contextmanager
def running_time(text):
    push(2)
    try:
        yield
    finally:
        msg(text + ' took {time} seconds.', 1)
        pop(2)
