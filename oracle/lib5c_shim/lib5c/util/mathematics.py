from oracle.thirdparty import gmean  # noqa: F401
