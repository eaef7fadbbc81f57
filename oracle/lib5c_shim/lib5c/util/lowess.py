from oracle.thirdparty import lowess  # noqa: F401
