"""Load the UNMODIFIED reference implementation for fixture generation.

Runs ONLY in the build container (needs /root/reference, which does not exist on
the GPU box).  Nothing under tests/ imports this at test time; it is used by
make_golden.py to produce the committed fixtures.

`import twoDSFS_class` fails (matplotlib/seaborn absent, module-level code opens
/Users/... paths: scripts/src/twoDSFS_class.py:12-16,1788-1790,1918), so we parse
the file, keep only the ClassDef (lines 20-1736) / the FunctionDefs of
scripts/sims_scan.py, and exec them in a namespace seeded with their imports.
No reference source is copied into this repository.
"""
import ast, gzip, os, glob, csv, math
import numpy as np
from scipy.stats import poisson, multinomial

REF = "/root/reference"


def _ns():
    return dict(gzip=gzip, os=os, glob=glob, csv=csv, math=math, np=np,
                poisson=poisson, multinomial=multinomial)


def load_class():
    path = f"{REF}/scripts/src/twoDSFS_class.py"
    tree = ast.parse(open(path).read())
    body = [n for n in tree.body if isinstance(n, ast.ClassDef)]
    ns = _ns()
    exec(compile(ast.Module(body=body, type_ignores=[]), path, "exec"), ns)
    return ns["LikelihoodInference_jointSFS"]


def load_legacy():
    """The first-generation script scripts/twoDSFS.py (Poisson composite score): its FunctionDefs only -- the module level
    opens /Users/... paths and imports matplotlib / pandas."""
    path = f"{REF}/scripts/twoDSFS.py"
    tree = ast.parse(open(path).read())
    want = {"calculate_2d_sfs", "normalize_2d_sfs", "calculate_p", "count_snps", "calculate_p_window"}
    body = [n for n in tree.body if isinstance(n, ast.FunctionDef) and n.name in want]
    ns = _ns()
    exec(compile(ast.Module(body=body, type_ignores=[]), path, "exec"), ns)
    return ns


def load_sims():
    path = f"{REF}/scripts/sims_scan.py"
    tree = ast.parse(open(path).read())
    body = [n for n in tree.body if isinstance(n, ast.FunctionDef)]
    ns = _ns()
    exec(compile(ast.Module(body=body, type_ignores=[]), path, "exec"), ns)
    return ns
