"""Host-side preparation of the streamflow-nudging inputs: which gage observations and last-observation rows belong
to the routed segments, and where they sit (compute.py:49-140 of the reference, restated on pandas indexes)."""
import pandas as pd


def prep_da_dataframes(usgs_df, lastobs_df, param_idx, exclude_segments=None):
    """-> (usgs_df_sub, lastobs_df_sub, positions of the gage segments in param_idx).

    Four cases (compute.py:64-72): observations + last-obs (analysis), last-obs only (forecast), observations only
    (cold start), neither (open loop).  Off-network upstream rows never count as gage locations (:80-83)."""
    subnet = param_idx.difference(list(exclude_segments)) if exclude_segments else param_idx
    have_obs = usgs_df is not None and not usgs_df.empty
    have_last = lastobs_df is not None and not lastobs_df.empty
    if have_obs and have_last:
        last_segs = lastobs_df.index.intersection(subnet).to_list()
        lastobs_sub = lastobs_df.loc[last_segs]
        usgs_segs = usgs_df.index.intersection(subnet).reindex(last_segs)[0].to_list()
        return usgs_df.loc[usgs_segs], lastobs_sub, param_idx.get_indexer(usgs_segs)
    if have_last:
        last_segs = lastobs_df.index.intersection(subnet).to_list()
        lastobs_sub = lastobs_df.loc[last_segs]
        # zero observation columns: the kernel then persists the last observation from step 1 (:104-108)
        return pd.DataFrame(index=lastobs_sub.index, columns=[]), lastobs_sub, param_idx.get_indexer(last_segs)
    if have_obs:
        usgs_segs = list(usgs_df.index.intersection(subnet))
        usgs_sub = usgs_df.loc[usgs_segs]
        return (usgs_sub, pd.DataFrame(index=usgs_sub.index, columns=["discharge", "time", "model_discharge"]),
                param_idx.get_indexer(usgs_segs))
    return pd.DataFrame(), pd.DataFrame(), []


def prep_da_positions_byreach(reach_list, gage_index):
    """-> (index of every reach that holds a gage, index of that gage in gage_index), in reach order (:124-140)."""
    members = set(gage_index)
    reach_key, reach_gage = [], []
    for i, reach in enumerate(reach_list):
        for s in reach:
            if s in members:
                reach_key.append(i)
                reach_gage.append(s)
    return reach_key, gage_index.get_indexer(reach_gage)
