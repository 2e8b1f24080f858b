from .utils import (
    convert_coo_to_csr,
    convert_coo_to_csr_indices_values,
    sparse_block_diag,
    sparse_block_diag_split,
    stack_csr,
)

__all__ = [
    "convert_coo_to_csr",
    "convert_coo_to_csr_indices_values",
    "sparse_block_diag",
    "sparse_block_diag_split",
    "stack_csr",
]
