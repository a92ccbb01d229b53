"""TEST INFRASTRUCTURE - not product code.  Only tests/, __graft_entry__ and bench.py's
reference / cpu_baseline legs may use what this produces.

Compiles the UNMODIFIED reference (woodywff/nas_3d_unet) from the sources where they lie under
/root/reference into CPython bytecode under oracle/_ref/ - the Python counterpart of compiling a C
reference into oracle/_ref/*.so: no reference source is copied into the repository, only the
compiler's output lands in the git-ignored oracle/_ref/, which travels to the GPU box with the
snapshot (like our own built .so) so that

  * bench.py --impl reference can time the REAL reference modules on the box's host cores
    (cpu_baseline.kind = "reference"), and
  * tests/test_gpu_drivers.py can run the reference's unchanged search.py / train.py /
    prediction.py against this package (nas_3d_unet_b200/dropin first on sys.path).

    python oracle/build_ref.py            # no-op (exit 0) when /root/reference is absent

The artefacts are stored as `<module>.bytecode` (snapshot tools tend to drop `*.pyc`);
reference_dir() materialises them as `<module>.pyc` in a per-process temporary directory, from which
CPython imports them like any sourceless module (`sys.path.insert(0, reference_dir())`).  The
bytecode is tied to the interpreter's minor version; the GPU box runs this same image.
"""
import atexit
import os
import py_compile
import shutil
import sys
import tempfile

HERE = os.path.dirname(os.path.abspath(__file__))
REF_SRC = os.environ.get("NAS3D_REFERENCE_DIR", "/root/reference")
REF_OUT = os.path.join(HERE, "_ref")


def build(verbose=True):
    if not os.path.isdir(REF_SRC):
        if verbose:
            print("oracle/build_ref: %s absent - keeping whatever oracle/_ref already holds" % REF_SRC)
        return None
    os.makedirs(REF_OUT, exist_ok=True)
    built = []
    for name in sorted(os.listdir(REF_SRC)):
        if not name.endswith(".py"):
            continue
        out = os.path.join(REF_OUT, name[:-3] + ".bytecode")
        py_compile.compile(os.path.join(REF_SRC, name), cfile=out, doraise=True, optimize=0)
        built.append(out)
    for stale in os.listdir(REF_OUT):
        if stale.endswith(".pyc"):
            os.remove(os.path.join(REF_OUT, stale))
    with open(os.path.join(REF_OUT, "PYTHON_VERSION"), "w") as f:
        f.write("%d.%d\n" % sys.version_info[:2])
    if verbose:
        print("oracle/build_ref: %d modules compiled into %s" % (len(built), REF_OUT))
    return REF_OUT


def reference_dir():
    """directory to put on sys.path to import the reference's modules: the sources where they lie
    (this container) or the compiled copy (GPU box); None if neither exists / version mismatch"""
    if os.path.isdir(REF_SRC) and os.path.exists(os.path.join(REF_SRC, "prim_ops.py")):
        return REF_SRC
    global _materialised
    if _materialised is not None:
        return _materialised
    ver = os.path.join(REF_OUT, "PYTHON_VERSION")
    if not (os.path.exists(ver) and os.path.exists(os.path.join(REF_OUT, "prim_ops.bytecode"))):
        return None
    with open(ver) as f:
        if f.read().strip() != "%d.%d" % sys.version_info[:2]:
            return None
    d = tempfile.mkdtemp(prefix="nas3d_ref_")
    atexit.register(shutil.rmtree, d, ignore_errors=True)
    for name in os.listdir(REF_OUT):
        if name.endswith(".bytecode"):
            shutil.copyfile(os.path.join(REF_OUT, name), os.path.join(d, name[:-len(".bytecode")] + ".pyc"))
    _materialised = d
    return d


_materialised = None


if __name__ == "__main__":
    build()
