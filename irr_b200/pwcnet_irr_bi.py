"""Drop-in for the reference's ``models/pwcnet_irr_bi.py`` ``PWCNet`` (bi-directional PWC-Net + IRR): same constructor, parameter names and
``forward({'input1','input2'}) -> {'flow'}`` (pwcnet_irr_bi.py:43-111).  The forward is shared: irr_b200/pwc_family.py."""
from .pwc_family import PWCFamily


class PWCNet(PWCFamily):
    IRR, BI, OCC = True, True, False
