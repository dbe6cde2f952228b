// hbt_b200.h - the one declaration a maintainer adds to HBT+ to use the batched post-unbinding properties.
//
// src/subhalo_tracking.cpp:901-906 of the reference reads
//     #pragma omp parallel for if(ParallelizeHaloes)
//     for(HBTInt i=0;i<Subhalos.size();i++)
//     {
//       Subhalos[i].CalculateProfileProperties(*this);
//       Subhalos[i].CalculateShape();
//     }
// and becomes
//     HBT_B200_CalculateProperties(Subhalos, *this);
// (defined in integration/subhalo_unbind_b200.cpp on top of hbtu_profile_batch, include/hbt_unbind.h).
#pragma once
#include "subhalo.h"
void HBT_B200_CalculateProperties(SubhaloList_t &Subhalos, const Snapshot_t &epoch);

// src/subhalo_tracking.cpp:531 of the reference reads `MaskSubhalos();` (inside SubhaloSnapshot_t::PrepareCentrals) and becomes
//     HBT_B200_MaskSubhalos(*this);
// (exclusive particle ownership on the device through hbtu_mask_batch; Subhalos / MemberTable are public members).
void HBT_B200_MaskSubhalos(SubhaloSnapshot_t &snap);

// src/subhalo_merge.cpp:193-198 of the reference (inside SubhaloSnapshot_t::MergeSubhalos): FillHelpers + DetectTraps become
//     std::vector<char> merged; HBT_B200_DetectTraps(*this, merged);  for(i...) Helpers[i].IsMerged = merged[i];
// (20-particle core moments, host chain walk and sink test on the device through hbtu_detect_traps).
void HBT_B200_DetectTraps(SubhaloSnapshot_t &snap, std::vector<char> &is_merged);

// src/subhalo_merge.cpp:207-214 of the reference (the two loops `if(Helpers[subid].IsMerged) Subhalos[subid].Unbind(*this);` and
// `... TruncateSource();`) become
//     HBT_B200_UnbindMerged(*this, merged);      // merged[subid] = Helpers[subid].IsMerged
// ONE batch for all merged hosts.  Without this patch the unmodified loop still works: the replaced Subhalo_t::Unbind is
// thread-safe and combines the concurrent OpenMP callers into batches (UnbindCombiner in subhalo_unbind_b200.cpp).
void HBT_B200_UnbindMerged(SubhaloSnapshot_t &snap, const std::vector<char> &is_merged);
