"""Drop-in for the reference's ``models/pwcnet_irr.py`` ``PWCNet`` (PWC-Net with iterative residual refinement): same constructor, parameter names and
``forward({'input1','input2'}) -> {'flow'}`` (pwcnet_irr.py:43-97).  The forward is shared: irr_b200/pwc_family.py."""
from .pwc_family import PWCFamily


class PWCNet(PWCFamily):
    IRR, BI, OCC = True, False, False
