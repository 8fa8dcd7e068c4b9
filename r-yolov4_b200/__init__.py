# This directory is imported under the name `ryolo_b200` (see ../ryolo_b200/__init__.py).
