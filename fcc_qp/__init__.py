"""Drop-in alias: ``from fcc_qp import FCCQP, FCCQPSolution, FCCQPOptions`` (the reference's
``fcc_qp/__init__.py:1`` does ``from fcc_qp_solver import *``) resolves to the B200 build."""
from fcc_qp_b200.fcc_qp_solver import *  # noqa: F401,F403
from fcc_qp_b200.fcc_qp_solver import FCCQP, FCCQPDetails, FCCQPOptions, FCCQPSolution  # noqa: F401
