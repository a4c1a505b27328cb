from relativepose_b200.model.mymodel import SCNet, weights_init  # noqa: F401
