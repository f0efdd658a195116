"""src/memory/statistics/mod.rs — same names, argument meaning and results; bodies are C-ABI calls."""
from __future__ import annotations

from ..anndata import IMAnnData
from ..shared import Direction


def compute_number(adata: IMAnnData, direction: Direction):
    """memory/statistics/mod.rs:10-15 -> Vec<u32>."""
    return adata.x().number(int(direction))


def compute_sum(adata: IMAnnData, direction: Direction):
    """:17-22 -> Vec<f64>."""
    return adata.x().sum(int(direction))


def compute_variance(adata: IMAnnData, direction: Direction):
    """:24-29 -> Vec<f64> (nonzero-only variance; helper/csr.rs:149-188)."""
    return adata.x().variance(int(direction))


def compute_min_max(adata: IMAnnData, direction: Direction):
    """:31-39 -> (Vec<f64>, Vec<f64>)."""
    return adata.x().min_max(int(direction))


def compute_std_dev(adata: IMAnnData, direction: Direction):
    """:41-46."""
    return adata.x().std_dev(int(direction))


class StatisticsContainer(dict):
    """memory/statistics/structs/mod.rs:1-10 (field names kept)."""
    __getattr__ = dict.__getitem__


def compute_qc_variables(adata: IMAnnData) -> StatisticsContainer:
    """:48-72 — one ABI call (two passes over the matrix instead of the reference's sixteen)."""
    q = adata.x().qc_all()
    return StatisticsContainer(num_per_cell=q["num_per_cell"], num_per_gene=q["num_per_gene"],
                               expr_per_gene=q["expr_per_gene"], expr_per_cell=q["expr_per_cell"],
                               variance_per_gene=q["variance_per_gene"], variance_per_cell=q["variance_per_cell"],
                               std_dev_per_cell=q["std_dev_per_cell"], std_dev_per_gene=q["std_dev_per_gene"])


def qc_vars_inplace(adata: IMAnnData) -> None:
    """:74-103 — column names exactly as the reference writes them into obs / var."""
    d = compute_qc_variables(adata)
    adata.obs["num_genes_per_cell"] = d.num_per_cell
    adata.obs["sum_expr_per_cell"] = d.expr_per_cell
    adata.obs["var_expr_per_cell"] = d.variance_per_cell
    adata.obs["std_dev_per_cell"] = d.std_dev_per_cell
    adata.var["num_cells_per_gene"] = d.num_per_gene
    adata.var["sum_expr_per_gene"] = d.expr_per_gene
    adata.var["var_expr_per_gene"] = d.variance_per_gene
    adata.var["std_dev_per_gene"] = d.std_dev_per_gene
