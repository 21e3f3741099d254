"""Drop-in for the reference's ``models/pwcnet_occ.py`` ``PWCNet`` (PWC-Net with an occlusion branch): same constructor, parameter names and
``forward({'input1','input2'}) -> {'flow','occ'}`` (pwcnet_occ.py:49-117).  The forward is shared: irr_b200/pwc_family.py."""
from .pwc_family import PWCFamily


class PWCNet(PWCFamily):
    IRR, BI, OCC = False, False, True
