from .Constants import *
from .potential_manager import *
from .imp_samp_manager import *
from .imp_samp import *
from .tensorflow_descriptors import *
from .sim_logger import *
from .file_manager import *
from .sim_archive import *
