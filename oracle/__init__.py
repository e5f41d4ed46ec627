"""ORACLE package: CPU restatements of the reference path.  Test infrastructure only -
the product (cm3d_b200/) never imports anything from here."""
