"""Drop-in for the reference's ``models/pwcnet_irr_occ.py`` ``PWCNet`` (PWC-Net + IRR with occlusion): same constructor, parameter names and
``forward({'input1','input2'}) -> {'flow','occ'}`` (pwcnet_irr_occ.py:47-112).  The forward is shared: irr_b200/pwc_family.py."""
from .pwc_family import PWCFamily


class PWCNet(PWCFamily):
    IRR, BI, OCC = True, False, True
