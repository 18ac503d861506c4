"""Env classes with device programs.  Each module here pairs with csrc/fam_<name>.cu."""
