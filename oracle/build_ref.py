"""Recipe: build the REAL reference's hot-path modules into oracle/_ref/ (git-ignored, shipped to the GPU box).

The reference is a pure-Python program, so "compiling it from the sources where they lie" means
byte-compiling: each module below is compiled by `py_compile` straight from /root/reference/src/<name>.py into a
SOURCELESS oracle/_ref/src/<name>.bytecode (the same interpreter runs here and on the GPU box: one image; the
extension is not `.pyc` because snapshot tools tend to drop `*.pyc`). No reference source text is copied into the
repository; oracle/_ref/ is listed in .gitignore like any other build output.

  python oracle/build_ref.py            # run by __graft_entry__.build() whenever /root/reference is present

What gets built: src/__init__, src/model, src/model_with_l2, src/train_utils, src/tempo_data, src/tempo_data_with_l2
-- everything `Trainer.train_step` (src/train_utils.py:149-183), `get_model` (src/model.py:708-759) and
`VAEWithL2Supervision.compute_loss` (src/model_with_l2.py:95-182) need. src/train_utils.py imports matplotlib
(:12-14), which this image does not have and which the train step never calls: a stub package (written by this script,
our own six lines) stands in for it.

Users: tests/ (oracle pinning on the GPU box, where /root/reference does not exist) and bench.py's reference arm
(`--impl reference`, cpu_baseline kind "reference"). The product package never imports it.
"""
import os
import py_compile
import sys

HERE = os.path.dirname(os.path.abspath(__file__))
REF = os.environ.get("TVAE_REFERENCE_ROOT", "/root/reference")
OUT = os.path.join(HERE, "_ref")
MODULES = ["__init__", "model", "model_with_l2", "train_utils", "tempo_data", "tempo_data_with_l2"]
EXT = ".bytecode"

MPL_STUB = '''"""Stand-in for matplotlib (absent from this image): the reference's train step imports it but never draws."""
def use(*a, **k):
    return None
'''
PYPLOT_STUB = '''def __getattr__(name):
    raise RuntimeError("matplotlib is stubbed out in oracle/_ref (plotting is outside the measured path)")
'''


def build(verbose=True):
    src = os.path.join(REF, "src")
    if not os.path.isdir(src):
        if verbose:
            print(f"build_ref: {src} not present, nothing built")
        return False
    os.makedirs(os.path.join(OUT, "src"), exist_ok=True)
    for m in MODULES:
        py_compile.compile(os.path.join(src, m + ".py"), cfile=os.path.join(OUT, "src", m + EXT), doraise=True,
                           dfile=f"<reference>/src/{m}.py")
    os.makedirs(os.path.join(OUT, "matplotlib"), exist_ok=True)
    with open(os.path.join(OUT, "matplotlib", "__init__.py"), "w") as f:
        f.write(MPL_STUB)
    with open(os.path.join(OUT, "matplotlib", "pyplot.py"), "w") as f:
        f.write(PYPLOT_STUB)
    with open(os.path.join(OUT, "BUILT_FROM"), "w") as f:
        f.write(f"{REF}/src  python {sys.version.split()[0]}\n")
    if verbose:
        print(f"build_ref: {len(MODULES)} modules byte-compiled into {OUT}/src")
    return True


def is_built():
    return os.path.exists(os.path.join(OUT, "src", "model" + EXT))


def load():
    """Import the built reference: returns the `src` package (src.model, src.train_utils, ...) or None when
    oracle/_ref has not been built. The stub matplotlib is only put on sys.path if the real one is missing. A process
    that already imported `src` from /root/reference itself (the CPU tests do) keeps using that one: same reference."""
    if "src.model" in sys.modules and "src.train_utils" in sys.modules:
        return sys.modules["src"]
    if not is_built():
        return None
    import importlib.machinery
    import importlib.util
    import types
    if OUT not in sys.path:
        try:
            import matplotlib  # noqa: F401
        except Exception:  # noqa: BLE001
            sys.path.insert(0, OUT)
    pkg = sys.modules.get("src")
    if pkg is None:
        pkg = types.ModuleType("src")
        pkg.__path__ = [os.path.join(OUT, "src")]
        sys.modules["src"] = pkg
    for m in MODULES[1:]:
        name = f"src.{m}"
        if name in sys.modules:
            continue
        loader = importlib.machinery.SourcelessFileLoader(name, os.path.join(OUT, "src", m + EXT))
        spec = importlib.util.spec_from_loader(name, loader)
        mod = importlib.util.module_from_spec(spec)
        sys.modules[name] = mod
        loader.exec_module(mod)
        setattr(pkg, m, mod)
    return pkg


if __name__ == "__main__":
    ok = build()
    sys.exit(0 if ok else 1)
