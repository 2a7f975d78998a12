#!/usr/bin/env python
"""Run the reference's UNMODIFIED geometric/scripts/main.py against this repo's B200 renderer:

    python scripts/run_reference_main.py --reference /path/to/3D-SDN/geometric -- --do train --dataset vkitti --mode full ...

Everything after `--` goes to the reference's absl flags.  See sdn_b200/compat.py for the shims that are installed."""
import argparse
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, os.path.join(ROOT, '3d-sdn_b200'))


def main():
    argv = sys.argv[1:]
    rest = []
    if '--' in argv:
        i = argv.index('--')
        argv, rest = argv[:i], argv[i + 1:]
    ap = argparse.ArgumentParser()
    ap.add_argument('--reference', default=None, help="the reference's geometric/ directory")
    args = ap.parse_args(argv)
    import sdn_b200.compat as compat
    compat.run_main(rest, args.reference)


if __name__ == '__main__':
    main()
