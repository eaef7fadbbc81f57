"""hic3defdr_b200: B200-native drop-in for the run_to_qvalues path of
thomasgilgenast/hic3defdr."""
__all__ = ['HiC3DeFDR']


def __getattr__(name):
    if name == 'HiC3DeFDR':
        from hic3defdr_b200.analysis import HiC3DeFDR
        return HiC3DeFDR
    raise AttributeError(name)
