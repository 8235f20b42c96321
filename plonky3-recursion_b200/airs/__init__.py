"""Table AIRs of the recursion circuit (host-side constraint + trace definitions).

Each module restates one reference AIR as an `eval(builder)` over the symbolic AirBuilder plus the matching
trace/preprocessed builders: Const/Public (`witness_send`), ALU (`alu`), Recompose (`recompose`), Poseidon2 (`poseidon2`).
"""
