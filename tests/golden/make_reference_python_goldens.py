"""Golden vectors from the reference's OWN Python for the pure-numpy helpers either side of the hot path (SURVEY.md
8(f)-3/4).  The reference package cannot be imported (tensorflow 1.14 / phi are absent), so the function sources are
extracted with `ast` from the files where they lie under /root/reference and executed unmodified against numpy
(np.float/np.int aliases restored for numpy >= 1.24).  Run in the build container only:

    python tests/golden/make_reference_python_goldens.py

Writes tests/golden/ref_python/*.npz|json; the tests read only those files."""
import ast
import json
import os
import sys
from collections.abc import Iterable

import numpy as np

REF = "/root/reference/diffpiso"
OUT = os.path.join(os.path.dirname(os.path.abspath(__file__)), "ref_python")


def extract(path, names):
    src = open(path).read()
    tree = ast.parse(src)
    ns = {"np": np, "Iterable": Iterable}
    if not hasattr(np, "float"):
        np.float = float
    if not hasattr(np, "int"):
        np.int = int
    for node in tree.body:
        if isinstance(node, ast.FunctionDef) and node.name in names:
            exec(compile(ast.Module([node], []), path, "exec"), ns)
    return [ns[n] for n in names]


def main():
    os.makedirs(OUT, exist_ok=True)
    (ek,) = extract(os.path.join(REF, "evaluation_tools.py"), ["EK_spectrum_2D"])
    rng = np.random.RandomState(7)
    for tag, shape in (("16x16", (16, 16)), ("12x20", (12, 20)), ("9x14", (9, 14))):
        field = rng.randn(shape[0], shape[1], 2).astype(np.float32)
        k, e = ek(field, [2 * np.pi, 2 * np.pi])
        np.savez(os.path.join(OUT, "ek_spectrum_%s.npz" % tag), field=field, k=k, e=e)
    assembler, load = extract(os.path.join(REF, "datamanagement.py"), ["data_path_assembler", "load_function"])
    args = dict(paths=["a/", "b/"], field_names=["velocity", "pressure"],
                characteristics=[[(0.1 * i, 0.05, 0.05) for i in range(12)], 3.0],
                start_frame=[2, 0], frame_count=[12, 7], step_count=[3, 2], dt_ratio=2)
    out = assembler(**args)
    json.dump(dict(args=args, out=out), open(os.path.join(OUT, "data_path_assembler.json"), "w"))
    print("written", sorted(os.listdir(OUT)))


if __name__ == "__main__":
    sys.exit(main())
