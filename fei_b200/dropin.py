"""Patch an importable reference `memdir_tools` package in place so its callers (fei.tools.*, the Flask server,
the CLIs) run the hot path on the GPU without any source change:

    import fei_b200.dropin; fei_b200.dropin.install()

replaces  memdir_tools.search.search_memories          (search.py:337)
          memdir_tools.utils.search_memories           (utils.py:299, the legacy substring search)
          memdir_tools.filter.FilterManager.process_memories / run_filters   (filter.py:188, :311)
          memdir_tools.filter.MemoryFilter.matches     (filter.py:67)
          memdir_tools.memorychain.MemoryChain.validate_chain                (memorychain.py:596)
and adds  memdir_tools.filter.apply_filters.  Names other modules already imported (`from memdir_tools.search import
search_memories` in server.py / cli.py, the package's own re-exports) are rebound too, by identity.
Everything else of the reference (SearchQuery objects, MemoryFilter objects, block classes, servers) is used as is:
the GPU layer only reads their public attributes."""
from __future__ import annotations

import importlib
import sys
from typing import Any


def _rebind(orig: Any, new: Any) -> None:
    """Point every module-level name that still holds `orig` (a `from x import f` done before install) at `new`."""
    for mod in list(sys.modules.values()):
        d = getattr(mod, "__dict__", None)
        if not d:
            continue
        for name, val in list(d.items()):
            if val is orig:
                d[name] = new


def install(package: str = "memdir_tools") -> None:
    ref_utils = importlib.import_module(f"{package}.utils")
    ref_search = importlib.import_module(f"{package}.search")
    ref_filter = importlib.import_module(f"{package}.filter")
    from .memdir_tools import filter as gfilter, search as gsearch, utils as gutils

    def _sync_base() -> None:
        gutils.set_memdir_base(ref_utils.MEMDIR_BASE)        # the tree the reference would walk (utils.py:16)

    def search_memories(query, folders=None, statuses=None, debug=False):
        _sync_base()
        return gsearch.search_memories(query, folders, statuses, debug)

    def _gpu_manager(mgr: Any) -> "gfilter.FilterManager":
        out = gfilter.FilterManager()
        for f in mgr.filters:
            g = gfilter.MemoryFilter(f.name)
            g.conditions, g.actions = f.conditions, f.actions
            out.add_filter(g)
        return out

    def process_memories(self, folders=None, statuses=None, dry_run=False):
        _sync_base()
        return _gpu_manager(self).process_memories(folders, statuses, dry_run)

    def matches(self, memory):
        g = gfilter.MemoryFilter(self.name)
        g.conditions = self.conditions
        return g.matches(memory)

    def run_filters(dry_run=False):
        _sync_base()
        return gfilter.run_filters(dry_run)

    def apply_filters(filters=None, folders=None, statuses=None, dry_run=False):
        _sync_base()
        if filters is not None and not isinstance(filters, gfilter.FilterManager) and hasattr(filters, "filters"):
            filters = _gpu_manager(filters)
        return gfilter.apply_filters(filters, folders, statuses, dry_run)

    def legacy_search_memories(query, folders=None, statuses=None, headers_only=False):
        _sync_base()
        return gutils.search_memories(query, folders, statuses, headers_only)

    _rebind(ref_search.search_memories, search_memories)
    ref_search.search_memories = search_memories
    _rebind(ref_utils.search_memories, legacy_search_memories)
    ref_utils.search_memories = legacy_search_memories
    ref_filter.FilterManager.process_memories = process_memories
    ref_filter.MemoryFilter.matches = matches
    _rebind(ref_filter.run_filters, run_filters)
    ref_filter.run_filters = run_filters
    ref_filter.apply_filters = apply_filters
    try:
        ref_chain = importlib.import_module(f"{package}.memorychain")
    except Exception:                                        # the reference module needs `requests`
        ref_chain = None
    if ref_chain is not None:
        from .memdir_tools.memorychain import install as install_chain
        install_chain(ref_chain)
