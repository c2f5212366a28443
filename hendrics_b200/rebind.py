"""Rebinding shim: run HENDRICS' own ``HENzsearch`` / ``HENefsearch`` on the B200 searches.

``hendrics/efsearch.py`` takes its two search functions from stingray (lines 15-18) and defines ``search_with_qffa``
(:928), ``search_with_ffa`` (:1092), ``dyn_folding_search`` (:1196) and ``folding_orbital_search`` (:263) itself; its
``folding_search`` (:1141) and ``_common_main`` (:1771) look all of them up as module globals at call time.  Replacing
those globals is therefore the whole integration: the reference's CLI, file naming, plotting and candidate handling run
unchanged, only the arithmetic moves to the GPU.

    import hendrics_b200.rebind as rb
    rb.install()                       # patch hendrics.efsearch in place (raises ImportError without HENDRICS)
    from hendrics.efsearch import main_zsearch
    main_zsearch(["events_ev.nc", "-f", "9.85", "-F", "9.95", "-N", "2", "--fast"])
    rb.uninstall()                     # put the stingray / in-tree functions back

or, as a process-wide switch, ``python -m hendrics_b200.rebind HENzsearch <args...>``.

``install(module)`` accepts any module object with those attributes (the tests pass a stand-in), so the shim itself is
checked without HENDRICS in the container.
"""

from __future__ import annotations

import importlib
import sys

# attribute of hendrics.efsearch -> (module of this package, attribute)
REBOUND = {
    "z_n_search": ("hendrics_b200.search", "z_n_search"),
    "epoch_folding_search": ("hendrics_b200.search", "epoch_folding_search"),
    "search_with_qffa": ("hendrics_b200.efsearch", "search_with_qffa"),
    "search_with_qffa_step": ("hendrics_b200.efsearch", "search_with_qffa_step"),
    "search_with_ffa": ("hendrics_b200.efsearch", "search_with_ffa"),
    "folding_orbital_search": ("hendrics_b200.orbital", "folding_orbital_search"),
}
# Not rebound on purpose: folding_search / _common_main / dyn_folding_search / transient_search / analyze_qffa_results
# stay the reference's own code (they reach the kernels through the names above; dyn_folding_search and the transient
# scan draw figures, which this package does not do).

_saved: dict = {}


def install(module=None):
    """Point the search functions of ``hendrics.efsearch`` (or of ``module``) at this package.  Returns the patched
    module.  Idempotent."""
    if module is None:
        module = importlib.import_module("hendrics.efsearch")
    saved = _saved.setdefault(id(module), {"module": module, "attrs": {}})
    for name, (modname, attr) in REBOUND.items():
        ours = getattr(importlib.import_module(modname), attr)
        if name not in saved["attrs"]:
            saved["attrs"][name] = getattr(module, name, None)
        setattr(module, name, ours)
    return module


def uninstall(module=None):
    """Undo :func:`install` (for ``module``, or for everything patched so far)."""
    for key in [k for k, v in _saved.items() if module is None or v["module"] is module]:
        entry = _saved.pop(key)
        for name, original in entry["attrs"].items():
            if original is None:
                delattr(entry["module"], name)
            else:
                setattr(entry["module"], name, original)


def is_installed(module) -> bool:
    return id(module) in _saved


def main(argv=None):
    argv = list(sys.argv[1:] if argv is None else argv)
    if not argv or argv[0] not in ("HENzsearch", "HENefsearch"):
        print(__doc__)
        return 2
    module = install()
    entry = module.main_zsearch if argv[0] == "HENzsearch" else module.main_efsearch
    entry(argv[1:])
    return 0


if __name__ == "__main__":
    raise SystemExit(main())
