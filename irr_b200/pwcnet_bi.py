"""Drop-in for the reference's ``models/pwcnet_bi.py`` ``PWCNet`` (bi-directional PWC-Net): same constructor, parameter names and
``forward({'input1','input2'}) -> {'flow'}`` (pwcnet_bi.py:41-109).  The forward is shared: irr_b200/pwc_family.py."""
from .pwc_family import PWCFamily


class PWCNet(PWCFamily):
    IRR, BI, OCC = False, True, False
