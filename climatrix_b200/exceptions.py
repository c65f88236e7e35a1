"""Exception names the hot path raises -- the reference's names and base classes.

Reference: src/climatrix/exceptions.py:5 (NotImplementedError subclass), :35
(ValueError subclass), :13 (KeyError subclass).  When the real climatrix package
is importable its classes are re-used, so ``except climatrix.exceptions.X``
written against the reference keeps catching what this package raises.
"""

try:  # pragma: no cover - climatrix is not installable in the build image
    from climatrix.exceptions import (  # type: ignore
        AxisMatchingError,
        DatasetCreationError,
        MissingAxisError,
        OperationNotSupportedForDynamicDatasetError,
        ReconstructorConfigurationFailed,
    )
except Exception:  # noqa: BLE001

    class DatasetCreationError(ValueError):
        """The input object cannot be interpreted as a climatrix dataset."""

    class OperationNotSupportedForDynamicDatasetError(NotImplementedError):
        """The operation is undefined for datasets with a non-trivial extra dimension."""

    class MissingAxisError(KeyError):
        """A required axis (latitude, longitude, point ...) is absent."""

    class AxisMatchingError(ValueError):
        """A coordinate name does not match any known axis."""

    class ReconstructorConfigurationFailed(ValueError):
        """A reconstructor received an inconsistent set of hyper-parameters."""


__all__ = [
    "AxisMatchingError",
    "DatasetCreationError",
    "MissingAxisError",
    "OperationNotSupportedForDynamicDatasetError",
    "ReconstructorConfigurationFailed",
]
