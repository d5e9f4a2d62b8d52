"""Locate and import Cirq (the library this backend plugs into).

Cirq itself is a dependency of the backend, not part of it: circuits, gates,
sweeps and the simulator driver classes (``cirq.SimulatorBase`` and friends,
reference ``cirq-core/cirq/sim/simulator_base.py:46``) are imported unchanged.

Search order: an already importable ``cirq``; ``$CIRQ_B200_CIRQ_PATH``;
``<repo>/baseline/_ref`` (``pip install --target`` of the unmodified
reference, travels to the GPU box); ``/root/reference/cirq-core`` (build
container only).

Two optional third-party modules that Cirq imports at module top for plotting
(``matplotlib``) and async fan-out of sampler jobs (``duet``) are absent from
this image.  Neither touches simulation arithmetic, so inert stand-ins are
registered when (and only when) the real modules cannot be imported.
"""
from __future__ import annotations

import asyncio
import functools
import importlib
import importlib.abc
import importlib.machinery
import importlib.util
import os
import sys
import types

_REPO_ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


class _Dummy:
    """Attribute sink used by the matplotlib stand-in."""

    def __init__(self, *args, **kwargs):
        pass

    def __call__(self, *args, **kwargs):
        return _Dummy()

    def __getattr__(self, name):
        if name.startswith('__') and name.endswith('__'):
            raise AttributeError(name)
        return _Dummy()

    def __iter__(self):
        return iter(())

    def __mro_entries__(self, bases):
        return (object,)


class _StubModule(types.ModuleType):
    def __getattr__(self, name):
        if name.startswith('__') and name.endswith('__'):
            raise AttributeError(name)
        return _Dummy()


class _StubFinder(importlib.abc.MetaPathFinder, importlib.abc.Loader):
    """Serves ``matplotlib[.*]`` / ``mpl_toolkits[.*]`` as inert packages."""

    def __init__(self, roots):
        self._roots = tuple(roots)

    def find_spec(self, fullname, path=None, target=None):
        root = fullname.split('.')[0]
        if root in self._roots:
            return importlib.machinery.ModuleSpec(fullname, self, is_package=True)
        return None

    def create_module(self, spec):
        mod = _StubModule(spec.name)
        mod.__path__ = []
        mod.__b200_stub__ = True
        return mod

    def exec_module(self, module):
        pass


def _install_matplotlib_stub():
    try:
        importlib.import_module('matplotlib')
        return
    except ImportError:
        pass
    sys.meta_path.append(_StubFinder(['matplotlib', 'mpl_toolkits']))


def _make_duet_stub():
    duet = types.ModuleType('duet')
    duet.__b200_stub__ = True

    def run(func, *args, **kwargs):
        return asyncio.run(func(*args, **kwargs))

    def sync(func):
        @functools.wraps(func)
        def wrapper(*args, **kwargs):
            return asyncio.run(func(*args, **kwargs))

        return wrapper

    async def pstarmap_async(func, iterable, *args, **kwargs):
        return [await func(*a) for a in iterable]

    async def pmap_async(func, iterable, *args, **kwargs):
        return [await func(a) for a in iterable]

    class Limiter:
        def __init__(self, capacity=None):
            self.capacity = capacity

        async def __aenter__(self):
            return self

        async def __aexit__(self, *exc):
            return False

    class AwaitableFuture(asyncio.Future):
        pass

    class AsyncCollector:
        def __init__(self):
            self._items = []
            self._done = False

        def add(self, item):
            self._items.append(item)

        def done(self):
            self._done = True

        def __aiter__(self):
            return self

        async def __anext__(self):
            if self._items:
                return self._items.pop(0)
            raise StopAsyncIteration

    class _Scope:
        def spawn(self, func, *args, **kwargs):
            return asyncio.ensure_future(func(*args, **kwargs))

    class new_scope:
        async def __aenter__(self):
            return _Scope()

        async def __aexit__(self, *exc):
            return False

    duet.run = run
    duet.sync = sync
    duet.pstarmap_async = pstarmap_async
    duet.pmap_async = pmap_async
    duet.Limiter = Limiter
    duet.AwaitableFuture = AwaitableFuture
    duet.AsyncCollector = AsyncCollector
    duet.new_scope = new_scope
    return duet


def _install_duet_stub():
    try:
        importlib.import_module('duet')
    except ImportError:
        sys.modules['duet'] = _make_duet_stub()


def candidate_paths():
    paths = []
    env = os.environ.get('CIRQ_B200_CIRQ_PATH')
    if env:
        paths.append(env)
    paths.append(os.path.join(_REPO_ROOT, 'baseline', '_ref'))
    paths.append('/root/reference/cirq-core')
    return paths


_cirq = None


def import_cirq():
    """Returns the ``cirq`` module, raising ImportError with guidance if absent."""
    global _cirq
    if _cirq is not None:
        return _cirq
    if 'cirq' in sys.modules:
        _cirq = sys.modules['cirq']
        return _cirq
    if importlib.util.find_spec('cirq') is None:
        for p in candidate_paths():
            if os.path.isdir(os.path.join(p, 'cirq')):
                sys.path.append(p)
                break
        else:
            raise ImportError(
                'cirq is not importable; install cirq-core or run '
                '`python -m pip install --no-deps --target baseline/_ref <cirq-core>` '
                '(see DESIGN.md).'
            )
    _install_matplotlib_stub()
    _install_duet_stub()
    _cirq = importlib.import_module('cirq')
    return _cirq


def cirq_available() -> bool:
    try:
        import_cirq()
        return True
    except ImportError:
        return False
