from .flags import FLAGS, FlagRegister
from .log import Log, CHECK, CHECK_NOTNULL, FatalError
from .dashboard import Dashboard, Monitor, monitor
from .timer import Timer

__all__ = ["FLAGS", "FlagRegister", "Log", "CHECK", "CHECK_NOTNULL", "FatalError", "Dashboard",
           "Monitor", "monitor", "Timer"]
