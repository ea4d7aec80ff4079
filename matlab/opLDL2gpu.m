classdef opLDL2gpu < handle
%OPLDL2GPU  GPU operator with the contract of ops/opLDL2.m: M*z solves
%   [A B'; B C] y = z  through device-resident LDL' factors, with the same public
%   properties (nitref, itref_tol, force_itref, residual_update; opLDL2.m:45-50).
%   It lets the UNMODIFIED kernels/cp*.m of the reference run with the
%   preconditioner apply on the GPU (e.g. when A is a matrix-free Spot operator).
    properties (SetAccess = private)
        h; n; nA; nC
    end
    properties
        nitref = 3; itref_tol = 1.0e-8; force_itref = false; residual_update = false
    end
    methods
        function op = opLDL2gpu(A, B, C)
            if nargin ~= 3, error('Invalid number of arguments.'); end          % opLDL2.m:61-63
            [L, D, P] = ldl([A B'; B C]);                                       % opLDL2.m:81-82
            op.h = cpk_b200_mex('ldl2_create', sparse(A), sparse(B), sparse(C), L, D, sparse(P));
            op.nA = size(A,1); op.nC = size(C,1); op.n = op.nA + op.nC;
        end
        function set.nitref(op, v), op.nitref = max(0, round(v)); op.push('nitref', op.nitref); end
        function set.itref_tol(op, v), op.itref_tol = v; op.push('itref_tol', v); end
        function set.force_itref(op, v), op.force_itref = logical(v); op.push('force_itref', double(v)); end
        function set.residual_update(op, v), op.residual_update = logical(v); op.push('residual_update', double(v)); end
        function y = mtimes(op, z), y = cpk_b200_mex('ldl2_apply', op.h, z); end             % opLDL2.m:161-188
        function x = mldivide(op, b), x = cpk_b200_mex('ldl2_matvec', op.h, b); end          % opLDL2.m:193-195
        function varargout = size(op, varargin), [varargout{1:nargout}] = size(sparse(op.n, op.n), varargin{:}); end
        function o = transpose(op), o = op; end                                            % opLDL2.m:120-136
        function o = ctranspose(op), o = op; end
        function delete(op), if ~isempty(op.h), cpk_b200_mex('destroy', op.h); end, end
    end
    methods (Access = private)
        function push(op, name, v)
            if ~isempty(op.h), cpk_b200_mex('ldl2_set', op.h, name, v); end
        end
    end
end
