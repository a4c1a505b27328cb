from relativepose_b200.model.mymodel import SCNet, Resnet18_8s, segmentation_layer, weights_init  # noqa: F401
