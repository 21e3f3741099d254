"""Drop-in for the reference's ``models/pwcnet_occ_bi.py`` ``PWCNet`` (bi-directional PWC-Net with occlusion): same constructor, parameter names and
``forward({'input1','input2'}) -> {'flow','occ'}`` (pwcnet_occ_bi.py:49-132).  The forward is shared: irr_b200/pwc_family.py."""
from .pwc_family import PWCFamily


class PWCNet(PWCFamily):
    IRR, BI, OCC = False, True, True
