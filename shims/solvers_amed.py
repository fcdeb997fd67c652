"""Import shim: lets the reference's `import solvers_amed` resolve to the B200-native implementation."""
from diff_sampler_b200.solvers_amed import *          # noqa: F401,F403
from diff_sampler_b200 import solvers_amed as _impl

globals().update({k: v for k, v in vars(_impl).items() if not k.startswith('__')})
