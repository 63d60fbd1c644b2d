# abundance_quantification_b200.R -- R side of the B200 binding (not runnable in this repository's image: no R here).
#
# Drop-in replacements, same contracts as the functions they replace in bambu 3.3.5:
#   abundance_quantification_b200()  for abundance_quantification()  R/bambu-quantify_utilityFunctions.R:379-391
#   bambu.quantDT.b200()             for bambu.quantDT()             R/bambu-quantify.R:53-74
# The C side is r/b200_shim.c (.Call entries over include/bambu_em.h).

#' All gene groups of one sample in ONE native call.
#' @param readClassDt the long table after filterTxRc / assignGroups, BEFORE split() (R/bambu-quantify.R:61-62):
#'   columns gene_grp_id, gene_sid, eqClassId, txid, aval, fullAval, uniqueAval, K, n.obs
#' @return numeric matrix with positional columns counts, fullLengthCounts, uniqueCounts, txid -- one row per
#'   (group, txid), what modifyQuantOut() consumes (R/bambu-quantify_utilityFunctions.R:438-443)
#' @noRd
abundance_quantification_b200 <- function(readClassDt, maxiter = 20000, conv = 10^(-2), minvalue = 10^(-8)) {
    dt <- copy(readClassDt)
    # rows of a group = its txids ascending; columns = its (gene_sid, eqClassId) ascending (getInputList, :359-371)
    rows <- unique(dt[, .(gene_grp_id, txid)])[order(gene_grp_id, txid)]
    cols <- dt[, .(K = K[1], n.obs = n.obs[1], multi = as.integer(any(uniqueAval == 0 & aval != 0))),
               by = .(gene_grp_id, gene_sid, eqClassId)][order(gene_grp_id, gene_sid, eqClassId)]
    rows[, r := .I - 1L]
    cols[, cg := .I - 1L]
    cols[, cl := cg - min(cg), by = gene_grp_id]          # column index local to its group
    dt <- rows[dt, on = c("gene_grp_id", "txid")]
    dt <- cols[, .(gene_grp_id, gene_sid, eqClassId, cl)][dt, on = c("gene_grp_id", "gene_sid", "eqClassId")]
    setorder(dt, r, cl)
    grp_row_ptr <- c(0, cumsum(rows[, .N, by = gene_grp_id]$N))
    grp_col_ptr <- c(0, cumsum(cols[, .N, by = gene_grp_id]$N))
    row_nnz_ptr <- c(0, cumsum(tabulate(dt$r + 1L, nbins = nrow(rows))))
    # "bambu_b200_em_batched_wire" takes the same arguments and sends the compact wire form (36 % of the bytes) to the GPU
    res <- .Call("bambu_b200_em_batched",
                 as.numeric(grp_row_ptr), as.numeric(grp_col_ptr), as.numeric(row_nnz_ptr),   # int64 travel as double
                 as.integer(dt$cl), as.numeric(dt$aval), as.raw(dt$fullAval != 0),
                 as.numeric(cols$n.obs), as.numeric(cols$K), as.raw(cols$multi),
                 as.integer(maxiter), as.numeric(minvalue), as.numeric(conv))
    # res$est is 3 x nrow(rows); res$iters / res$status are per group (the reference drops `iter`, src/em.cpp:93-94).
    # The single-transcript shortcut of run_parallel (:410-414) is applied inside the library.
    cbind(t(res$est), rows$txid)
}

#' bambu.quantDT() with the packing, the EM and the merge inside the library (device packer, csrc/quant_dev.cu).
#' @noRd
bambu.quantDT.b200 <- function(readClassDt, emParameters = NULL, verbose = FALSE) {
    emParameters <- setEmParameters(emParameters)                      # R/bambu_utilityFunctions.R:48-53
    res <- .Call("bambu_b200_quantDT",
                 match(readClassDt$GENEID, unique(readClassDt$GENEID)),              # simplifyNames' gene_sid (:236-241)
                 as.integer(readClassDt$txid), as.integer(readClassDt$eqClassId), as.numeric(readClassDt$nobs),
                 as.logical(readClassDt$equal), as.numeric(readClassDt$rcWidth), as.numeric(readClassDt$txlen),
                 as.logical(emParameters$degradationBias), as.integer(emParameters$maxiter),
                 as.numeric(emParameters$minvalue), as.numeric(emParameters$conv))
    data.table(txid = res$txid, counts = res$counts, fullLengthCounts = res$fullLengthCounts,
               uniqueCounts = res$uniqueCounts)
}

# In R/bambu-quantify.R:64-70 either swap
#     outEst <- abundance_quantification(inputRcDt, readClassDt, ...)          # after split(), per group
# for
#     outEst <- abundance_quantification_b200(readClassDt, ...)               # before split(), all groups at once
# or replace bambu.quantDT by bambu.quantDT.b200.  A CUDA context does not survive fork(): run the quantification
# stage with ncore = 1 per GPU process (R/bambu.R:187-191), or hoist the per-sample loop into one
# bambu_quantDT_cohort call (include/bambu_em.h) from the master process.
.onUnload_b200 <- function(libpath) .Call("bambu_b200_shutdown")
