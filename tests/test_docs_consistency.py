"""CPU: the documents name files that exist — every `tests/…`, `tools/…`, `profiles/…`, `examples/…`, `oracle/…`,
`include/…`, `diffusion-pipe_b200/…` path written in DESIGN.md, INTEGRATION.md, README.md and profiles/README.md, and every
`file.py::test_name` test id."""
import os
import re

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
DOCS = ['DESIGN.md', 'INTEGRATION.md', 'README.md', os.path.join('profiles', 'README.md')]
PREFIXES = ('tests/', 'tools/', 'profiles/', 'examples/', 'oracle/', 'include/', 'diffusion-pipe_b200/')


def test_paths_and_test_ids_in_the_documents_exist():
    missing = []
    for doc in DOCS:
        text = open(os.path.join(ROOT, doc)).read()
        for m in re.finditer(r'`([A-Za-z0-9_./\-]+(?:::[A-Za-z0-9_\[\]\-]+)?)`', text):
            tok = m.group(1)
            path, _, test = tok.partition('::')
            if not path.startswith(PREFIXES) or '*' in path or path.endswith('/'):
                continue
            full = os.path.join(ROOT, path)
            if not os.path.exists(full):
                missing.append((doc, tok))
            elif test and not re.search(r'def ' + re.escape(test.split('[')[0]) + r'\b', open(full).read()):
                missing.append((doc, tok))
    assert not missing, missing


def test_product_code_never_imports_the_oracle_or_the_tests():
    """the oracle is the checker: nothing under diffusion-pipe_b200/ (nor train.py) may import `oracle`, `tests`, the kernel
    test doubles or a golden helper; bench.py and __graft_entry__.py may, in their baseline / smoke legs only"""
    import ast
    banned = {'oracle', 'tests', 'kernel_doubles', 'synth', 'toy_model'}
    offenders = []
    files = [os.path.join(ROOT, 'train.py')]
    for d, _, fs in os.walk(os.path.join(ROOT, 'diffusion-pipe_b200')):
        files += [os.path.join(d, f) for f in fs if f.endswith('.py')]
    for path in files:
        for node in ast.walk(ast.parse(open(path).read())):
            names = []
            if isinstance(node, ast.Import):
                names = [a.name for a in node.names]
            elif isinstance(node, ast.ImportFrom) and node.level == 0:
                names = [node.module or '']
            for n in names:
                if n.split('.')[0] in banned:
                    offenders.append((os.path.relpath(path, ROOT), n))
    assert not offenders, offenders
    # bench.py: the oracle appears only inside the baseline helpers
    tree = ast.parse(open(os.path.join(ROOT, 'bench.py')).read())
    for fn in tree.body:
        if isinstance(fn, ast.FunctionDef):
            uses = any(isinstance(n, ast.ImportFrom) and (n.module or '').split('.')[0] == 'oracle' for n in ast.walk(fn))
            assert uses == (fn.name in ('cpu_reference_sample', 'gpu_library_sample')), fn.name
    assert not any(isinstance(n, (ast.Import, ast.ImportFrom)) and 'oracle' in ast.dump(n) for n in tree.body)
