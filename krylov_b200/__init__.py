"""krylov_b200 -- B200-native (sm_100a) implementation of the iteration hot
path of ju-liu/krylov: ``cg`` / ``minres`` / ``gmres`` with the reference's call
signatures, on hand-written CUDA kernels behind a C ABI (include/krylov_b200.h).

Importing this package loads ``csrc/libkrylov_b200.so`` and raises if it is
missing; there is no CPU fallback.
"""
from ._lib import LIB_PATH, KrylovB200Error  # noqa: F401  (loads the library)
from .errors import ArgumentError  # noqa: F401
from .csr import CsrMatrix  # noqa: F401
from .operators import (  # noqa: F401
    Identity,
    Info,
    LinearOperatorWrapper,
    Product,
    aslinearoperator,
    get_default_inner,
)
from .cg import cg  # noqa: F401
from .minres import minres  # noqa: F401
from .gmres import gmres  # noqa: F401
from .shortrec import (  # noqa: F401
    bicg, bicgstab, cgne, cgnr, cgr, cgs, chebyshev, gcr, qmr, symmlq)
from .givens import givens  # noqa: F401
from .householder import Householder  # noqa: F401
from .arnoldi import ArnoldiHouseholder, ArnoldiLanczos, ArnoldiMGS  # noqa: F401
from . import stencils  # noqa: F401
from . import utils  # noqa: F401

__version__ = "0.1.0"
