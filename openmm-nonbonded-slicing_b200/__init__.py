"""B200-native SlicedNonbondedForce hot path (host-side mirror of the plugin's operator API).

Import with ``importlib.import_module("openmm-nonbonded-slicing_b200")`` (the directory name is the
one the task fixes; it is not a valid identifier).  The CUDA library ``csrc/libnbslice_b200.so`` is
loaded lazily when the first kernel is created and there is no CPU fallback.
"""
from .api import (ONE_4PI_EPS0, B200CalcSlicedNonbondedForceKernel, CalcSlicedNonbondedForceKernel, Context,
                  OpenMMException, Platform, SlicedKernelBase, SlicedNonbondedForce, SlicedNonbondedForceImpl,
                  State, System, build_desc, sliceIndex)
from . import abi

__all__ = ["ONE_4PI_EPS0", "B200CalcSlicedNonbondedForceKernel", "CalcSlicedNonbondedForceKernel", "Context",
           "OpenMMException", "Platform", "SlicedKernelBase", "SlicedNonbondedForce", "SlicedNonbondedForceImpl",
           "State", "System", "build_desc", "sliceIndex", "abi"]
