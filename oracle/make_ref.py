#!/usr/bin/env python3
"""
oracle/make_ref.py -- TEST INFRASTRUCTURE.  Stages the UNMODIFIED reference scripts of the hot path under
oracle/_ref/ so that the CPU arm of bench.py (`--impl reference`, `cpu_baseline`) can time the reference's own code
on the GPU box, where /root/reference does not exist.

The reference is pure Python (no build step), so "compiling it into oracle/_ref" is a byte-for-byte copy of the files
the path needs, straight from where they lie under /root/reference, each recorded with its sha256 in MANIFEST.json.
oracle/_ref/ is git-ignored (reference sources never enter this repository's history) but travels to the GPU box with
the working tree, like the built .so files.  Nothing in the product imports it; bench.py falls back to the oracle port
(`kind: "port"`) when the directory is absent.

    python oracle/make_ref.py            # run in the build container; __graft_entry__.build() does it too
"""
import hashlib
import json
import os
import shutil
import sys

REF = '/root/reference'
HERE = os.path.dirname(os.path.abspath(__file__))
DEST = os.path.join(HERE, '_ref')
# the model classes, the driver that owns the bootstrap loop, the model-table generator, and the licence
FILES = ('HOMOGENOUES_MODELS.py', 'RECENT_ADMIXTURE_MODELS.py', 'DISTANT_ADMIXTURE_MODELS.py', 'ANEUPLOIDY_TEST.py',
         'MAKE_STATISTICAL_MODEL.py', 'LICENSE')


def stage(verbose=True):
    """ Copies the files and writes the manifest; returns False when the reference is not mounted. """
    if not os.path.isdir(REF):
        if verbose:
            print('make_ref: %s is not present, nothing staged' % REF)
        return False
    os.makedirs(DEST, exist_ok=True)
    manifest = {}
    for name in FILES:
        src, dst = os.path.join(REF, name), os.path.join(DEST, name)
        shutil.copyfile(src, dst)
        with open(dst, 'rb') as f:
            manifest[name] = hashlib.sha256(f.read()).hexdigest()
    with open(os.path.join(DEST, 'MANIFEST.json'), 'w') as f:
        json.dump({'source': REF, 'note': 'unmodified copies, see oracle/make_ref.py', 'sha256': manifest}, f, indent=1)
    if verbose:
        print('make_ref: staged %d files under %s' % (len(FILES), DEST))
    return True


def available():
    """ True when every staged file is present and matches its recorded digest. """
    try:
        with open(os.path.join(DEST, 'MANIFEST.json')) as f:
            manifest = json.load(f)['sha256']
        for name, digest in manifest.items():
            with open(os.path.join(DEST, name), 'rb') as f:
                if hashlib.sha256(f.read()).hexdigest() != digest:
                    return False
        return all(n in manifest for n in FILES)
    except Exception:
        return False


if __name__ == '__main__':
    sys.exit(0 if stage() else 1)
