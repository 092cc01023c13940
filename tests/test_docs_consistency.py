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
