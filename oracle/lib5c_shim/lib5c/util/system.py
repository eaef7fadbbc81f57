from oracle.thirdparty import check_outdir  # noqa: F401
