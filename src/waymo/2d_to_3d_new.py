"""README name of the lifting stage (`python 2d_to_3d_new.py`): same script as 2d_to_3d.py."""
import os
import runpy

if __name__ == "__main__":
    runpy.run_path(os.path.join(os.path.dirname(os.path.abspath(__file__)), "2d_to_3d.py"), run_name="__main__")
