"""Console entry points mirroring the reference's ``HENzsearch`` / ``HENefsearch`` scripts
(``pyproject.toml:105,130`` of HENDRICS):

    python -m hendrics_b200.cli HENzsearch events_ev.p -f 9.85 -F 9.95 -N 2 [--fast]
    python -m hendrics_b200.cli HENefsearch events_ev.p -f 9.85 -F 9.95 -n 64
"""

from __future__ import annotations

import sys


def main(argv=None):
    argv = list(sys.argv[1:] if argv is None else argv)
    if not argv or argv[0] not in ("HENzsearch", "HENefsearch"):
        print(__doc__)
        return 2
    from .efsearch import main_efsearch, main_zsearch

    fn = main_zsearch if argv[0] == "HENzsearch" else main_efsearch
    for out in fn(argv[1:]):
        print(out)
    return 0


if __name__ == "__main__":
    raise SystemExit(main())
