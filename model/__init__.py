"""Drop-in import name: ``from model.mymodel import SCNet`` keeps working (evaluation.py:16).  Implementation:
relativepose_b200/model/mymodel.py."""
