from ._lobpcg import lobpcg, lobpcg_func, BlockMatrices  # noqa: F401
