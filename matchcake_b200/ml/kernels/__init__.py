from .fermionic_pqc_kernel import FermionicPQCKernel
from .gram_matrix import GramMatrix
from .kernel import Kernel
from .linear_nif_kernel import LinearNIFKernel
from .nif_kernel import NIFKernel
from .nystroem import Nystroem

__all__ = ["FermionicPQCKernel", "GramMatrix", "Kernel", "LinearNIFKernel", "NIFKernel", "Nystroem"]
