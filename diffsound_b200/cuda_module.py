"""Loader of the native plugin, with the reference's names.

API mirror of src/cuda_module.py:7-41: `CUDA_MODULE.load(Debug, MemoryCheck,
Verbose)`, `CUDA_MODULE.get(name)` and the module-level `mass_matrix_assembler`.
The reference JIT-compiles a pybind11 torch extension (`diffFEM`) at import
time; here `load()` opens the prebuilt C-ABI library
`libdiffsound_sm100.so` through ctypes (no JIT, no pybind) and returns an object
whose attributes are callables taking torch tensors.  `assemble_mass_matrix`
keeps the reference signature (src/cuda/massMatrixDouble.h:14-15):

    assemble_mass_matrix(vertices, tets, values, rows, cols, element_mm, density, order)

with the same flat layouts and dtypes; it launches on torch's current stream
(the reference uses the legacy default stream, include/macro.h:149).
"""
from . import _lib, native


class _Module:
    """Attribute access like the reference's pybind module."""

    def __init__(self):
        self._lib = _lib.load()
        self.assemble_mass_matrix = native.assemble_mass_coo

    def __getattr__(self, name):
        # raw C-ABI entry points are reachable under their own names as well
        return getattr(self._lib, name)


class CUDA_MODULE:
    _module = None

    @staticmethod
    def get(name):
        if CUDA_MODULE._module is None:
            CUDA_MODULE.load()
        return getattr(CUDA_MODULE._module, name)

    @staticmethod
    def load(Debug=False, MemoryCheck=False, Verbose=False):
        """Debug / MemoryCheck selected nvcc flags of the reference's JIT build; the prebuilt library
        has one configuration (-O3 -lineinfo, argument checks always on), so they are accepted and
        ignored.  Verbose prints the library path and version."""
        CUDA_MODULE._module = _Module()
        if Verbose:
            print(f"diffsound_b200: loaded {_lib.LIB_PATH} (version {CUDA_MODULE._module._lib.ds_version()})")
        return CUDA_MODULE._module


CUDA_MODULE.load()
mass_matrix_assembler = CUDA_MODULE.get("assemble_mass_matrix")
