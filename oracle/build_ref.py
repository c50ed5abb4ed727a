"""Recipe that stages the UNMODIFIED reference orchestration layer under ``oracle/_ref/``.  TEST
INFRASTRUCTURE ONLY (see ``oracle/__init__.py``).

    python oracle/build_ref.py            # also run by __graft_entry__.build()

The reference is pure Python: there is nothing to compile.  What this recipe "builds" is an importable
tree ``oracle/_ref/edrgp/`` holding byte-identical copies of the five reference modules that make up
its orchestration layer L3 -- ``edrgp/{__init__,edr,base,utils,datasets}.py`` -- taken from where they
lie under ``/root/reference`` (the same files ``pip install --target`` would lay down; the reference's
own ``setup.py`` cannot run here), plus ``MANIFEST.json`` with their SHA-256
digests.  ``edrgp/gp_model`` is NOT staged: it needs GPy, which is not installable in this environment.

``oracle/_ref/`` is listed in ``.gitignore`` (no reference source ever enters the history) but not in
``.gpurunignore``: like the built ``.so`` it travels to the GPU box, where ``/root/reference`` does not
exist, so that the ``-m gpu`` tests can drive the CUDA estimator / transformer through the reference's
own ``EffectiveDimensionalityReduction`` (``edrgp/edr.py:11``, ``edrgp/base.py:435-466``) and compare
``refit`` / ``get_estimator_gradients`` (``edrgp/base.py:202-239``, ``edrgp/edr.py:199-241``) value by
value.  Nothing under ``edrgp_b200/`` imports it (``tests/test_abi.py``).
"""
import hashlib
import json
import os
import shutil
import sys

HERE = os.path.dirname(os.path.abspath(__file__))
REFERENCE = '/root/reference'
DEST = os.path.join(HERE, '_ref')
MODULES = ('__init__.py', 'edr.py', 'base.py', 'utils.py', 'datasets.py')


def _sha256(path):
    h = hashlib.sha256()
    with open(path, 'rb') as f:
        h.update(f.read())
    return h.hexdigest()


def available():
    """True when a staged copy is importable (this container after ``build``, or the GPU box)."""
    return all(os.path.exists(os.path.join(DEST, 'edrgp', f)) for f in MODULES)


def build(verbose=False):
    """Stage the copy when ``/root/reference`` is present; otherwise leave what is there untouched.
    Returns the directory to put on ``sys.path`` (or None when neither source nor copy exists)."""
    src = os.path.join(REFERENCE, 'edrgp')
    if not os.path.isdir(src):
        return DEST if available() else None
    out = os.path.join(DEST, 'edrgp')
    os.makedirs(out, exist_ok=True)
    manifest = {'source': src, 'files': {}}
    for f in MODULES:
        shutil.copyfile(os.path.join(src, f), os.path.join(out, f))
        manifest['files']['edrgp/' + f] = _sha256(os.path.join(out, f))
    with open(os.path.join(DEST, 'MANIFEST.json'), 'w') as fh:
        json.dump(manifest, fh, indent=1, sort_keys=True)
    if verbose:
        print("staged %d reference modules under %s" % (len(MODULES), out))
    return DEST


def verify():
    """The staged files still match their recorded digests (nobody edited the copy)."""
    with open(os.path.join(DEST, 'MANIFEST.json')) as fh:
        manifest = json.load(fh)
    return all(_sha256(os.path.join(DEST, rel)) == digest for rel, digest in manifest['files'].items())


def path():
    """Directory holding an importable unmodified ``edrgp`` (L3 only): the real reference when it is
    on this machine, else the staged copy, else None."""
    if os.path.isdir(os.path.join(REFERENCE, 'edrgp')):
        return REFERENCE
    return DEST if available() else None


if __name__ == '__main__':
    print(build(verbose=True))
    sys.exit(0 if available() and verify() else 1)
