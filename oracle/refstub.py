"""TEST INFRASTRUCTURE ONLY -- import the *unmodified* reference (dynast v1.0.1) in this container.

The reference is pure Python but depends on packages that are not installed here
(``ngs_tools``, ``pysam``, ``anndata``, ``pystan``).  This module registers minimal stand-ins for
them in ``sys.modules`` and puts ``/root/reference`` on ``sys.path`` so that
``dynast.preprocessing.{conversion,aggregation}`` and ``dynast.estimation.{p_e,p_c,alpha}`` import
and run unchanged (recipe: SURVEY.md Appendix B).

It is used ONLY by ``oracle/make_golden.py`` (to generate the committed vectors under
``tests/golden/``) and by the optional ``tests/test_oracle_vs_reference.py`` checks, which skip when
``/root/reference`` is absent (it does not exist on the GPU box).  Nothing under
``dynast-release_b200/`` may import this file.
"""
import gzip
import logging
import os
import pickle
import sys
import tempfile
import types

REFERENCE_ROOT = os.environ.get('DYNAST_REFERENCE_ROOT', '/root/reference')


def available() -> bool:
    return os.path.isdir(os.path.join(REFERENCE_ROOT, 'dynast'))


class _Anything:
    """Attribute-tolerant placeholder for modules whose contents are never exercised."""

    def __init__(self, *a, **k):
        pass

    def __getattr__(self, name):
        return _Anything()

    def __call__(self, *a, **k):
        return _Anything()


def _placeholder_module(name):
    mod = types.ModuleType(name)
    mod.__getattr__ = lambda attr: _Anything  # type: ignore
    return mod


class _Segment:
    """Shape-compatible with the pickled ``ngs_tools.gtf.Segment.Segment`` objects in genes.pkl.gz."""

    def __init__(self, start=0, end=0):
        self._start = start
        self._end = end

    @property
    def start(self):
        return self._start

    @property
    def end(self):
        return self._end

    def __iter__(self):
        return iter((self._start, self._end))

    def __repr__(self):
        return f'Segment({self._start}, {self._end})'


def _flatten_iter(it):
    for x in it:
        if isinstance(x, (list, tuple, set, frozenset)):
            yield from _flatten_iter(x)
        else:
            yield x


def _mkstemp(dir=None, delete=False):
    fd, path = tempfile.mkstemp(dir=dir)
    os.close(fd)
    if delete:
        os.remove(path)
    return path


def _write_pickle(obj, path, *args, **kwargs):
    with gzip.open(path, 'wb') as f:
        pickle.dump(obj, f, *args, **kwargs)
    return path


def _read_pickle(path):
    with gzip.open(path, 'rb') as f:
        return pickle.load(f)


def _merge_dictionaries(d1, d2, f=None, default=None):
    """ngs_tools.utils.merge_dictionaries: union of keys in first-appearance order (d1 then d2),
    leaves combined with ``f`` (operator.add by default), missing leaves use ``default`` or the type's
    zero value."""
    import operator
    f = f or operator.add

    def merge(a, b):
        out = {}
        keys = list(a.keys()) + [k for k in b.keys() if k not in a]
        for k in keys:
            va, vb = a.get(k), b.get(k)
            if isinstance(va, dict) or isinstance(vb, dict):
                out[k] = merge(va or {}, vb or {})
            else:
                if va is None:
                    va = default if default is not None else type(vb)()
                if vb is None:
                    vb = default if default is not None else type(va)()
                out[k] = f(va, vb)
        return out

    return merge(d1, d2)


def _flatten_dictionary(d, keys=None):
    keys = keys or ()
    for k, v in d.items():
        if isinstance(v, dict):
            yield from _flatten_dictionary(v, keys + (k,))
        else:
            yield keys + (k,), v


def install():
    """Idempotently install the stand-ins and make ``import dynast`` work. Returns the dynast package."""
    if not available():
        raise RuntimeError(f'reference not found at {REFERENCE_ROOT}')
    if 'ngs_tools' not in sys.modules:
        ngs = types.ModuleType('ngs_tools')
        ngs.__path__ = []

        class Logger(logging.Logger):

            def __init__(self, name='ref'):
                super().__init__(name)

            def namespaced(self, ns):
                return lambda fn: fn

            def namespaced_context(self, ns):
                import contextlib
                return contextlib.nullcontext()

        m_logging = types.ModuleType('ngs_tools.logging')
        m_logging.Logger = Logger
        m_logging.set_logger = lambda logger: None

        m_utils = types.ModuleType('ngs_tools.utils')
        m_utils.flatten_iter = _flatten_iter
        m_utils.mkstemp = _mkstemp
        m_utils.read_pickle = _read_pickle
        m_utils.write_pickle = _write_pickle
        m_utils.all_exists = lambda *paths: all(os.path.exists(p) for p in paths)
        m_utils.merge_dictionaries = _merge_dictionaries
        m_utils.flatten_dictionary = _flatten_dictionary
        for name in ('run_executable', 'open_as_text', 'decompress_gzip', 'flatten_dict_values'):
            setattr(m_utils, name, _Anything())

        m_seq = types.ModuleType('ngs_tools.sequence')
        _comp = str.maketrans('ACGTNacgtn', 'TGCANtgcan')
        m_seq.complement_sequence = lambda s, reverse=False: s.translate(_comp)[::-1 if reverse else 1]

        m_gtf = types.ModuleType('ngs_tools.gtf')
        m_gtf.__path__ = []
        m_gtf.Segment = _Segment
        m_gtf_seg = types.ModuleType('ngs_tools.gtf.Segment')
        m_gtf_seg.Segment = _Segment
        _Segment.__module__ = 'ngs_tools.gtf.Segment'

        ngs.logging, ngs.utils, ngs.sequence, ngs.gtf = m_logging, m_utils, m_seq, m_gtf
        sys.modules.update({
            'ngs_tools': ngs,
            'ngs_tools.logging': m_logging,
            'ngs_tools.utils': m_utils,
            'ngs_tools.sequence': m_seq,
            'ngs_tools.gtf': m_gtf,
            'ngs_tools.gtf.Segment': m_gtf_seg,
        })
        for sub in ('progress', 'bam', 'chemistry', 'fastq'):
            mod = _placeholder_module(f'ngs_tools.{sub}')
            setattr(ngs, sub, mod)
            sys.modules[f'ngs_tools.{sub}'] = mod
        for name in ('pysam', 'anndata', 'pystan'):
            if name not in sys.modules:
                sys.modules[name] = _placeholder_module(name)
    if REFERENCE_ROOT not in sys.path:
        sys.path.insert(0, REFERENCE_ROOT)
    import dynast  # noqa: E402
    import dynast.utils as dutils

    def _no_progress(counter, total, *async_results, desc=None):
        for r in async_results:
            r.wait()

    dutils.display_progress_with_counter = _no_progress
    dutils.as_completed_with_progress = lambda futures: __import__('concurrent.futures').futures.as_completed(futures)
    return dynast


def fixture(*parts):
    return os.path.join(REFERENCE_ROOT, 'tests', 'fixtures', *parts)
