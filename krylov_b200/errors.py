class ArgumentError(Exception):
    """Raised when an argument is invalid (same name and role as the
    reference's ``krylov.errors.ArgumentError``, errors.py:1-9): e.g. stepping an
    Arnoldi process whose Krylov subspace was already found invariant."""

    def __init__(self, message):
        super().__init__(message)
