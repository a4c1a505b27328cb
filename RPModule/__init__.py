"""Drop-in import name: ``from RPModule.rpmodule import RelativePoseEstimation_helper`` keeps working
(evaluation.py:13,16; trainRelativePoseModuleRecFD.py:6-7).  Implementation: relativepose_b200/RPModule."""
