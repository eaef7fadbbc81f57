from oracle.thirdparty import adjust_pvalues  # noqa: F401
