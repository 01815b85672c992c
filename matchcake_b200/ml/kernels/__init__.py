from .fermionic_pqc_kernel import FermionicPQCKernel
from .gram_matrix import GramMatrix
from .kernel import Kernel
from .nif_kernel import NIFKernel
from .nystroem import Nystroem

__all__ = ["FermionicPQCKernel", "GramMatrix", "Kernel", "NIFKernel", "Nystroem"]
