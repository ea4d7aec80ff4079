function [x, stats, flag] = reg_cpkrylov_gpu(method, b, A, B, C, G, opts)
%REG_CPKRYLOV_GPU  Drop-in for reg_cpkrylov (same signature, reg_cpkrylov.m:1) that
% runs the Krylov loop and every preconditioner apply on a B200 through
% cpk_b200_mex / libcpk_b200.so.  Only the factorization stays on the host
% (untimed setup, stats.ptime), exactly where the reference calls ldl (opLDL2.m:82).
%
% A may be a matrix or a linear operator (reg_cpkrylov.m:40): a Spot operator or a function
% handle is evaluated on the host, through mexCallMATLAB, from inside the one device launch
% (only v and A*v cross the bus per iteration); B, C and G must be explicit matrices.
    if (nargin < 6)
        error('reg_cpkrylov: not enough inputs');                 % reg_cpkrylov.m:122-125
    end
    if nargin < 7, opts = struct(); end
    ids = struct('cpcg',0,'cpcglanczos',1,'cpminres',2,'cpsymmlq',3,'cpgmres',4,'cpdqgmres',5);
    name = func2str(method);
    if ~isfield(ids, name), error('reg_cpkrylov_gpu: unknown method %s', name); end

    tstartp = tic;
    n = size(A,1); m = size(B,1);
    K = [G B'; B -C];
    [L, D, P] = ldl(K);                                           % opLDL2.m:81-82 (HSL MA57)
    hM = cpk_b200_mex('ldl2_create', sparse(G), sparse(B), sparse(-C), L, D, sparse(P));
    if isnumeric(A)
        hS = cpk_b200_mex('system_create', sparse(A), sparse(C), hM);
    else
        hS = cpk_b200_mex('system_create_op', n, A, sparse(C), hM);
    end
    ptime = toc(tstartp);
    cleanup = onCleanup(@() cpk_b200_mex('destroy', hS));

    names = {'nitref','itref_tol','residual_update','force_itref'};   % reg_cpkrylov.m:135-148
    for k = 1:numel(names)
        if isfield(opts, names{k}), cpk_b200_mex('ldl2_set', hM, names{k}, double(opts.(names{k}))); end
    end

    tstarts = tic;
    ov = nan(1,6); f = {'atol','rtol','btol','itmax','restart','mem'};
    for k = 1:6
        if isfield(opts, f{k}), ov(k) = double(opts.(f{k})); end
    end
    [x, niters, solved, status, hist, tdev] = cpk_b200_mex('reg_solve', hS, ids.(name), b(:), ov, [n m]);
    stats.niters = niters;
    stats.t_device = tdev;                                        % CUDA-event time of the one launch
    doprint = true;                                               % default of every solver (e.g. cpminres.m:110)
    if isfield(opts, 'print'), doprint = opts.print; end
    if doprint
        % the reference prints one line per iteration from inside the loop (cpminres.m:167-173,
        % 239-241); the device loop cannot, so the table is printed from the returned history
        fprintf('\n**** %s on B200 ****\n\n%5s  %9s\n', name, 'iter', '|resid|');
        for k = 1:size(hist,1), fprintf('%5d  %9.2e\n', k-1, hist(k,1)); end
        fprintf('\n');
    end
    if strcmp(name, 'cpsymmlq')                                   % cpsymmlq.m:363-366
        stats.cgresidHistory = hist(:,1); stats.lqresidHistory = hist(:,2); stats.qrresidHistory = hist(:,3);
    else
        stats.residHistory = hist(:,1);
    end
    if strcmp(name, 'cpcglanczos')                                % cpcglanczos.m:312-324
        txt = {'maximum number of iterations attained', 'residual small compared to initial residual', 'backward error small'};
        stats.status = txt{status+1};
    end
    flag.solved = solved;
    stats.ptime = ptime;                                          % reg_cpkrylov.m:177-178
    stats.stime = toc(tstarts);
end
