"""Crude undefined-name check (no pyflakes in this image): names that are loaded somewhere in a module but bound nowhere in it
(imports, assignments, defs, arguments, comprehension / except / with targets) and are not builtins.  Catches the
forgotten-import class of error in code paths that only run on a GPU box.  Usage: python tools/undef_check.py file.py ..."""
import ast
import builtins
import sys


def undefined_names(path):
    tree = ast.parse(open(path).read(), path)
    bound = set(dir(builtins)) | {"__file__", "__name__", "__doc__"}
    for n in ast.walk(tree):
        if isinstance(n, (ast.Import, ast.ImportFrom)):
            for a in n.names:
                bound.add((a.asname or a.name).split(".")[0])
        elif isinstance(n, (ast.FunctionDef, ast.AsyncFunctionDef, ast.ClassDef)):
            bound.add(n.name)
        elif isinstance(n, ast.arg):
            bound.add(n.arg)
        elif isinstance(n, ast.Name) and isinstance(n.ctx, (ast.Store, ast.Del)):
            bound.add(n.id)
        elif isinstance(n, ast.ExceptHandler) and n.name:
            bound.add(n.name)
        elif isinstance(n, (ast.Global, ast.Nonlocal)):
            bound.update(n.names)
    return sorted({(n.lineno, n.id) for n in ast.walk(tree)
                   if isinstance(n, ast.Name) and isinstance(n.ctx, ast.Load) and n.id not in bound})


if __name__ == "__main__":
    bad = 0
    for f in sys.argv[1:]:
        for line, name in undefined_names(f):
            print("%s:%d: undefined name %s" % (f, line, name))
            bad += 1
    sys.exit(1 if bad else 0)
