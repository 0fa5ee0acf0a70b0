// ref_sort.cpp -- TEST INFRASTRUCTURE ONLY. Runs the reference's five sort kernels (LocalRadixSort, PreScan, BlockSum,
// GlobalScan, GlobalRadixSort; text generated from Assets/_Shaders/Sorting/*.compute by build_ref.sh) under the
// lock-step wave emulator, sequenced as ComputeBufferSorter.Sort() sequences them (ComputeBufferSorter.cs:100-126).
#include "hlsl_shim.hpp"
#include "wave_emulator.hpp"

#include <cstring>
#include <vector>

namespace ref_local_radix_sort {
#include "../_ref/gen_LocalRadixSort.inc"
}
namespace ref_scan {
#include "../_ref/gen_Scan.inc"
}
namespace ref_global_radix_sort {
#include "../_ref/gen_GlobalRadixSort.inc"
}

// Constants.cginc:1-5 (macros in the generated text); the sort is hard-wired to 512 groups of 1024 elements
static const int kThreads = 1024, kBlocks = 512, kBucket = 256, kScanGroups = kBlocks * kBucket / kThreads /* 128 */;

static void k_local(uint32_t t, uint32_t g) { ref_local_radix_sort::LocalRadixSort(uint3(t, 0, 0), uint3(g, 0, 0)); }
static void k_prescan(uint32_t t, uint32_t g) { ref_scan::PreScan(uint3(t, 0, 0), uint3(g, 0, 0)); }
static void k_blocksum(uint32_t t, uint32_t g) { ref_scan::BlockSum(uint3(t, 0, 0), uint3(g, 0, 0)); }
static void k_globalscan(uint32_t t, uint32_t g) { ref_scan::GlobalScan(uint3(t, 0, 0), uint3(g, 0, 0)); }
static void k_global(uint32_t t, uint32_t g) { ref_global_radix_sort::GlobalRadixSort(uint3(t, 0, 0), uint3(g, 0, 0)); }

extern "C" {

// One pass of ComputeBufferSorter.Sort() (:104-116) over `groups` x 1024 elements (the reference always dispatches
// all 512 groups; groups that are not dispatched here are empty blocks: their columns of the digit-major count table
// are zero). keys / values: in and out (GlobalRadixSort writes back into them, :87-88). The intermediates come out in
// the reference's own layouts: offsets[block * 256 + digit], sizes[digit * 512 + block] before and after the scan.
void usrt_ref_sort_pass(uint* keys, uint* values, int groups, int bit_offset, uint* sorted_blocks_keys,
                        uint* sorted_blocks_values, uint* offsets /* 512*256 */, uint* sizes_before /* 512*256 */,
                        uint* sizes_after /* 512*256 */) {
    std::vector<uint> sizes((size_t)kBlocks * kBucket, 0u), block_sums(kScanGroups, 0u);
    std::memset(offsets, 0, sizeof(uint) * kBlocks * kBucket);
    {   // ComputeBufferSorter.cs:107 Dispatch(LocalRadixSort)
        using namespace ref_local_radix_sort;
        keysData.data = keys; valuesData.data = values;
        sortedBlocksKeysData.data = sorted_blocks_keys; sortedBlocksValuesData.data = sorted_blocks_values;
        offsetsData.data = offsets; sizesData.data = sizes.data();
        bitOffset = bit_offset;
        wave_emu::dispatch(k_local, kThreads, groups);
    }
    std::memcpy(sizes_before, sizes.data(), sizeof(uint) * sizes.size());
    {   // :112-114 PreScan (128 groups), BlockSum (1 group of 128 threads), GlobalScan (128 groups)
        using namespace ref_scan;
        data.data = sizes.data(); blockSumsData.data = block_sums.data();
        wave_emu::dispatch(k_prescan, kThreads, kScanGroups);
        wave_emu::dispatch(k_blocksum, kScanGroups, 1);
        wave_emu::dispatch(k_globalscan, kThreads, kScanGroups);
    }
    std::memcpy(sizes_after, sizes.data(), sizeof(uint) * sizes.size());
    {   // :116 Dispatch(GlobalRadixSort)
        using namespace ref_global_radix_sort;
        sortedBlocksKeysData.data = sorted_blocks_keys; sortedBlocksValuesData.data = sorted_blocks_values;
        offsetsData.data = offsets; sizesData.data = sizes.data();
        sortedKeysData.data = keys; sortedValuesData.data = values;
        bitOffset = bit_offset;
        wave_emu::dispatch(k_global, kThreads, groups);
    }
}

// ComputeBufferSorter.Sort() :100-126 -- bitOffset = 0, 8, 16, 24
void usrt_ref_sort(uint* keys, uint* values, int groups) {
    const size_t n = (size_t)groups * kThreads;
    std::vector<uint> sbk(n), sbv(n), off((size_t)kBlocks * kBucket), sb((size_t)kBlocks * kBucket), sa((size_t)kBlocks * kBucket);
    for (int bit_offset = 0; bit_offset < 32; bit_offset += 8)
        usrt_ref_sort_pass(keys, values, groups, bit_offset, sbk.data(), sbv.data(), off.data(), sb.data(), sa.data());
}

}  // extern "C"
