"""Static undefined-name check over every Python source of the repository: most device-engine / bench branches only run on a GPU
box, so a misspelt or out-of-scope name there would otherwise surface at round end, not here (no pyflakes in the image)."""
import ast
import builtins
import os

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
SKIP_DIRS = {".git", "baseline/_ref", "gpurun_out", ".work", "build", "__pycache__"}


def _sources():
    for d, dirs, files in os.walk(ROOT):
        rel = os.path.relpath(d, ROOT)
        dirs[:] = [x for x in dirs if x not in SKIP_DIRS and os.path.join(rel, x).lstrip("./") not in SKIP_DIRS]
        for f in files:
            if f.endswith(".py"):
                yield os.path.join(d, f)


class _Scope:
    def __init__(self, parent, kind):
        self.parent, self.kind, self.names, self.globals = parent, kind, set(), set()


class _Checker(ast.NodeVisitor):
    """Two passes per scope: collect every binding, then resolve every load (function bodies see module names bound later)."""

    def __init__(self):
        self.problems = []
        self.builtins = set(dir(builtins)) | {"__file__", "__name__", "__doc__", "__spec__", "__path__", "__class__"}

    # ---- binding collection
    def _bind_target(self, scope, t):
        for n in ast.walk(t):
            if isinstance(n, ast.Name) and isinstance(n.ctx, (ast.Store, ast.Del)):
                scope.names.add(n.id)

    def _collect(self, scope, body):
        for node in body:
            for n in self._walk_same_scope(node):
                if isinstance(n, (ast.FunctionDef, ast.AsyncFunctionDef, ast.ClassDef)):
                    scope.names.add(n.name)
                elif isinstance(n, (ast.Import, ast.ImportFrom)):
                    for a in n.names:
                        scope.names.add((a.asname or a.name).split(".")[0])
                        if a.name == "*":
                            scope.names.add("*")
                elif isinstance(n, ast.Name) and isinstance(n.ctx, (ast.Store, ast.Del)):
                    scope.names.add(n.id)
                elif isinstance(n, ast.ExceptHandler) and n.name:
                    scope.names.add(n.name)
                elif isinstance(n, (ast.Global, ast.Nonlocal)):
                    scope.globals.update(n.names)
                    scope.names.update(n.names)
                elif isinstance(n, ast.MatchAs) and n.name:
                    scope.names.add(n.name)

    def _walk_same_scope(self, node):
        """Nodes of ``node`` that belong to the current scope (nested function / class / comprehension bodies excluded; their
        names, decorators and defaults included)."""
        stack = [node]
        while stack:
            n = stack.pop()
            yield n
            if isinstance(n, (ast.FunctionDef, ast.AsyncFunctionDef, ast.Lambda)):
                if not isinstance(n, ast.Lambda):
                    stack.extend(n.decorator_list)
                    if n.returns:
                        stack.append(n.returns)
                a = n.args
                stack.extend(a.defaults + [d for d in a.kw_defaults if d is not None])
                stack.extend(x.annotation for x in a.args + a.posonlyargs + a.kwonlyargs + [y for y in (a.vararg, a.kwarg) if y]
                             if x.annotation is not None)
                continue
            if isinstance(n, ast.ClassDef):
                stack.extend(n.decorator_list + n.bases + [k.value for k in n.keywords])
                continue
            if isinstance(n, (ast.ListComp, ast.SetComp, ast.DictComp, ast.GeneratorExp)):
                stack.append(n.generators[0].iter)       # evaluated in the enclosing scope
                continue
            stack.extend(ast.iter_child_nodes(n))

    # ---- resolution
    def _defined(self, scope, name):
        s = scope
        first = True
        while s is not None:
            if (s.kind != "class" or first) and (name in s.names or "*" in s.names):
                return True
            first = False
            s = s.parent
        return name in self.builtins

    def run_scope(self, scope, body, where):
        self._collect(scope, body)
        for node in body:
            for n in self._walk_same_scope(node):
                if isinstance(n, ast.Name) and isinstance(n.ctx, ast.Load) and not self._defined(scope, n.id):
                    self.problems.append(f"{where}:{n.lineno}: undefined name {n.id!r}")
                elif isinstance(n, (ast.FunctionDef, ast.AsyncFunctionDef, ast.Lambda)):
                    sub = _Scope(scope, "function")
                    a = n.args
                    for x in a.args + a.posonlyargs + a.kwonlyargs + [y for y in (a.vararg, a.kwarg) if y]:
                        sub.names.add(x.arg)
                    self.run_scope(sub, n.body if not isinstance(n, ast.Lambda) else [ast.Expr(n.body)], where)
                elif isinstance(n, ast.ClassDef):
                    self.run_scope(_Scope(scope, "class"), n.body, where)
                elif isinstance(n, (ast.ListComp, ast.SetComp, ast.DictComp, ast.GeneratorExp)):
                    sub = _Scope(scope, "function")
                    for g in n.generators:
                        self._bind_target(sub, g.target)
                    inner = [g.iter for g in n.generators[1:]] + [c for g in n.generators for c in g.ifs]
                    inner += [n.key, n.value] if isinstance(n, ast.DictComp) else [n.elt]
                    self.run_scope(sub, [ast.Expr(e) for e in inner], where)


def check_file(path):
    with open(path) as f:
        src = f.read()
    tree = ast.parse(src, path)
    c = _Checker()
    c.run_scope(_Scope(None, "module"), tree.body, os.path.relpath(path, ROOT))
    return c.problems


def test_checker_finds_a_planted_mistake(tmp_path):
    p = tmp_path / "m.py"
    p.write_text("import os\n\ndef f(a):\n    def g():\n        return a + b_typo\n    return [os.sep for q in range(a) if q] + [late]\n\nlate = 1\n"
                 "class C:\n    x = 1\n    def m(self):\n        return x\n")
    probs = check_file(str(p))
    assert any("b_typo" in q for q in probs) and any("'x'" in q for q in probs) and len(probs) == 2, probs


def test_no_undefined_names_in_the_repository():
    problems = []
    for path in _sources():
        problems += check_file(path)
    assert not problems, "\n".join(problems)
