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

CPython imports `name.pyc` files found directly in a sys.path directory (sourceless import), so
`sys.path.insert(0, "oracle/_ref")` is all a user of the compiled reference needs.  The bytecode is
tied to the interpreter's minor version; the GPU box runs this same image.
"""
import os
import py_compile
import sys

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
        out = os.path.join(REF_OUT, name + "c")
        py_compile.compile(os.path.join(REF_SRC, name), cfile=out, doraise=True, optimize=0)
        built.append(out)
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
    ver = os.path.join(REF_OUT, "PYTHON_VERSION")
    if os.path.exists(ver) and os.path.exists(os.path.join(REF_OUT, "prim_ops.pyc")):
        with open(ver) as f:
            if f.read().strip() == "%d.%d" % sys.version_info[:2]:
                return REF_OUT
    return None


if __name__ == "__main__":
    build()
