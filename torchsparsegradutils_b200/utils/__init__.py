"""Index / layout helpers of the sparse_mm path (see utils.py): the subset of the reference's
``torchsparsegradutils.utils`` that ``sparse_mm`` stands on, backed by the sm_100a index-builder kernels."""
from . import utils as _impl

stack_csr = _impl.stack_csr
sparse_block_diag = _impl.sparse_block_diag
sparse_block_diag_split = _impl.sparse_block_diag_split
convert_coo_to_csr = _impl.convert_coo_to_csr
convert_coo_to_csr_indices_values = _impl.convert_coo_to_csr_indices_values

__all__ = [name for name in ("stack_csr", "sparse_block_diag", "sparse_block_diag_split", "convert_coo_to_csr",
                             "convert_coo_to_csr_indices_values")]
