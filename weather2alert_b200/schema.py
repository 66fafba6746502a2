"""Feature-store schema of the HeatAlertEnv hot path.

The column order is the one the reference produces when it merges its three
parquet files (`/root/reference/src/weather2alert/env.py:49-56`) whose columns are
written by the ETL (`/root/reference/data-processing/merge_state_actions.py:228-248`
for the exogenous block and `:264-272` for the endogenous/action block).  The
observation the reference returns is that merged row with three slots overwritten and
one label *appended* (`env.py:186-195`), hence 29 slots.
"""
from __future__ import annotations

EXOGENOUS = (
    "heat_qi",
    "heat_qi_above_25",
    "heat_qi_above_75",
    "hi_max",
    "hi_max_above_25",
    "hi_max_above_75",
    "hi_max*heat_qi",
    "hi_max_above_25*heat_qi",
    "hi_max_above_75*heat_qi",
    "heat_qi_3d",
    "excess_heat_3d",
    "excess_heat_3d*heat_qi",
    "heat_qi_7d",
    "excess_heat_7d",
    "excess_heat_7d*heat_qi",
    "weekend",
    "holiday",
    "dos",
    "bspline_dos_0",
    "bspline_dos_1",
    "bspline_dos_2",
)
ENDOGENOUS = (
    "alert",
    "alerts_2wks",
    "alert_lag1",
    "alert_streak",
    "remaining_budget",
    "issued_in_advance",
    "significance",
)

#: the 28 merged data columns, in merged-frame order (obs slots 0..27)
DATA_COLUMNS = EXOGENOUS + ENDOGENOUS
#: the 29 observation slots (`info["feature_names"]` of the reference)
OBS_NAMES = DATA_COLUMNS + ("alert_2wks",)
OBS_DIM = len(OBS_NAMES)  # 29

#: the 27 numeric reward features = every data column but `significance`
#: (`/root/reference/reward-training/modules.py:265,345-346`), slots 0..26
REWARD_FEATURES = DATA_COLUMNS[:27]
N_REWARD_FEATURES = 27
#: coefficient vector = 27 features + bias (`/root/reference/reward-training/train.py:121-126`)
N_COEF = 28

SLOT_HEAT_QI = 0
SLOT_ALERT_LAG1 = 23  # overwritten by the agent's last (clamped) action
SLOT_ALERT_STREAK = 24  # overwritten by the agent's streak
SLOT_REMAINING_BUDGET = 25  # overwritten by budget - alerts issued
SLOT_SIGNIFICANCE = 27  # non-numeric in the reference (str / NaN) -> float code here
SLOT_ALERT_2WKS = 28  # appended label: agent's 14-day alert count

#: slots whose value comes from the feature table when the reward is evaluated
TABLE_REWARD_SLOTS = tuple(range(0, 23)) + (26,)  # 24 slots

MAX_DAYS = 153  # May 1 .. Sep 30 (`data-processing/conf/config.yaml:6-9`)
REWARD_SCALE = -1000.0 / 152.0  # `env.py:221`
