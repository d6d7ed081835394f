"""Import alias: ``import seld_b200`` -> the package directory ``sound-event-localization-detection_b200``
(whose name, fixed by the project layout, is not a valid Python identifier)."""
import importlib
import sys

_pkg = importlib.import_module("sound-event-localization-detection_b200")
sys.modules[__name__] = _pkg
for _name, _mod in list(sys.modules.items()):
    if _name.startswith("sound-event-localization-detection_b200."):
        sys.modules["seld_b200." + _name.split(".", 1)[1]] = _mod
