"""Exception types of krylov_b200.

``ArgumentError`` keeps the name and role it has in the reference package
(``krylov.errors``): it is what stepping an Arnoldi/Lanczos process raises once
its Krylov subspace was found invariant, and what the solvers raise when they
would need such a step.  It derives from ``KrylovB200Error``'s sibling base so
callers can catch everything this package raises with one ``except`` clause.
"""


class KrylovError(Exception):
    """Base class of the algorithmic errors raised by this package."""


class ArgumentError(KrylovError):
    pass
