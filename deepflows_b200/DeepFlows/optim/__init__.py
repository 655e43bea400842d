from .optimier import Optimizer  # noqa: F401
from .adadelta import Adadelta  # noqa: F401
from .adagrad import Adagrad  # noqa: F401
from .adam import Adam  # noqa: F401
from .sgd import SGD  # noqa: F401
from . import scheduler  # noqa: F401
