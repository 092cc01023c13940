"""Import alias: the product package lives in the directory ``diffusion-pipe_b200/`` (the name the
task mandates, which is not a valid Python identifier).  This shim makes it importable as
``diffusion_pipe_b200`` by pointing ``__path__`` at that directory and running its ``__init__``."""
import os as _os

_real = _os.path.join(_os.path.dirname(_os.path.dirname(_os.path.abspath(__file__))), 'diffusion-pipe_b200')
__path__ = [_real]
with open(_os.path.join(_real, '__init__.py')) as _f:
    exec(compile(_f.read(), _os.path.join(_real, '__init__.py'), 'exec'))
del _f
